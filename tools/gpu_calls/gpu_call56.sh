#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call56
VPB_COOP_PDL=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "layernorm" 2>&1 | tail -3
for f in 0 1 0 1; do
  VPB_COOP_PDL=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.bench.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.bench.json').read().strip().splitlines()[-1])
print('coop_pdl=$f', round(r['value'],1), round(r['ms_per_step'],3), 'e2e', round(r['e2e']['value'],1))" || tail -3 $O.err.txt
done
