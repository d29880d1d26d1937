#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call8
nvidia-smi -L > $O.smi.txt
timeout 1500 python -m pytest tests -x -q -m gpu -s > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O.bench2.json 2> $O.bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --train > $O.train2.json 2> $O.train2.err
VPB_NCCL_MAX_CTAS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --train > $O.train2_defcta.json 2>> $O.train2.err
VPB_GRAD_BF16=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 --train > $O.train2_bf16.json 2>> $O.train2.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 3 --train > $O.train1.json 2>> $O.train2.err
grep -E "passed|failed|rc=|peaked|well-posed" $O.tests.txt | tail -30
for f in train1 train2 train2_defcta train2_bf16; do python -c "
import json
d=json.loads(open('$O.$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value']), round(d['ms_per_step'],3), d['clocks'])
"; done
python -c "
import json
d=json.loads(open('$O.bench2.json').read().strip().splitlines()[-1])
print('bench2', round(d['value']), d['ms_per_step'])
for k,v in d.get('configs',{}).items(): print(' ',k, round(v['value']), round(v['ms_per_step'],2))
"
tail -5 $O.bench2.err
