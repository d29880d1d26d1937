#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call39
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 2 --train --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.$name.json 2> $O.$name.err
  python -c "
import json
r=json.loads(open('$O.$name.json').read().strip().splitlines()[-1])
print('$name', round(r['value']), round(r['ms_per_step'],2), 'e2e ms', round(r['e2e']['ms_per_step'],2))"
}
run default A=1
run noprefetch VPB_BENCH_PREFETCH=0
run logdev VPB_LOG_ON_DEVICE=1
run logdev_noprefetch VPB_LOG_ON_DEVICE=1 VPB_BENCH_PREFETCH=0
