#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call45
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 4 --steps 10 --warmup 3 > $O.bench4.json 2> $O.bench4.err
python - <<PY
import json
d=json.loads(open('$O.bench4.json').read().strip().splitlines()[-1])
print('bench4', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))
for k,v in d.get('configs',{}).items(): print(' ',k, round(v['value']), round(v['ms_per_step'],2), round(v['e2e']['value']))
PY
