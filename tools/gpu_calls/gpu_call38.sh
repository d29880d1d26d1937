#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call38
for i in 1 2; do
timeout 300 python bench.py --train --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.train.json 2>$O.err.txt
python -c "
import json
r=json.loads(open('$O.train.json').read().strip().splitlines()[-1])
print('train', r['value'], r['ms_per_step'], r['e2e'])"
done
