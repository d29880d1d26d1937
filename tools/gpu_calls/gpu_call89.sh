#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call89
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra --no-cpu-baseline > $O.bench2.json 2> $O.bench2.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$O.bench2.json').read().strip().splitlines()[-1])
print('bench2', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['roofline']['ms_per_step_with_events'], d['n_gpus'])" || tail -5 $O.bench2.err
