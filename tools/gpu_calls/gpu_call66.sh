#!/bin/bash
# 2 GPUs: NCCL tests (incl. ViTPose+ ranks with different datasets), default bench under torchrun
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call66
nvidia-smi -L > $O.smi.txt
timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_moe.py -x -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?"; tail -15 $O.tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O.bench2.json 2> $O.bench2.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$O.bench2.json').read().strip().splitlines()[-1])
print('bench2', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))
for k,v in d.get('configs',{}).items(): print(' ',k, round(v['value']), round(v['ms_per_step'],2), 'e2e', round(v['e2e']['value']))
"
tail -3 $O.bench2.err
