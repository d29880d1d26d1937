#!/bin/bash
# 8 GPUs: training step scaling, default vs bf16 gradient exchange vs capped NCCL CTAs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call72
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 8 --steps 15 --warmup 3 --train --no-cpu-baseline > $O.$name.json 2> $O.$name.err
  python -c "
import json
d=json.loads(open('$O.$name.json').read().strip().splitlines()[-1])
print('$name', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['clocks'])
" || tail -5 $O.$name.err
}
run default VPB_DUMMY=1
run bf16grad VPB_GRAD_BF16=1
run cta16 VPB_NCCL_MAX_CTAS=16
