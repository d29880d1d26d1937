#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call70
for v in 0 0; do
VPB_LOG_ON_DEVICE=$v timeout 300 python bench.py --train --steps 20 --warmup 3 --no-cpu-baseline > $O.train.json 2>$O.err.txt
python -c "
import json
r=json.loads(open('$O.train.json').read().strip().splitlines()[-1])
print('log_on_device=$v train', round(r['value'],1), round(r['ms_per_step'],3), 'e2e', round(r['e2e']['ms_per_step'],3))" || tail -3 $O.err.txt
done
