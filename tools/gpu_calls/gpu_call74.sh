#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call74
for v in 64 32 48 24 64 32; do
  VPB_HOST_CHUNK=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.bench.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.bench.json').read().strip().splitlines()[-1])
print('host_chunk=$v', round(r['value'],1), round(r['ms_per_step'],3), 'e2e', round(r['e2e']['value'],1), round(r['e2e']['ms_per_step'],3), r['clocks']['sm_mhz'])" | tee -a $O.bench.txt
done
