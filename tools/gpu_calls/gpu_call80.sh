#!/bin/bash
# 4 GPUs, clean rebuild of the library: smoke, training / NCCL tests, default bench under torchrun
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call80
python -c "import __graft_entry__ as g; g.smoke()" > $O.smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $O.smoke.txt
timeout 600 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_train_step.py -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?"; tail -2 $O.tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 > $O.bench4.json 2> $O.bench4.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$O.bench4.json').read().strip().splitlines()[-1])
print('bench4', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['clocks'])
for k,v in d.get('configs',{}).items(): print(' ',k, round(v['value']), round(v['ms_per_step'],2), 'e2e', round(v['e2e']['value']))
"
