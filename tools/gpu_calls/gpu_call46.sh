#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call46
for f in 0 1 0 1; do
  echo "fullgrid=$f" >> $O.gemm.txt
  VPB_LN_FULLGRID=$f timeout 200 python tools/gemm_time.py 256 base 2>&1 | grep -E "^proj_ln|^fc2_ln" >> $O.gemm.txt
  VPB_LN_FULLGRID=$f timeout 200 python tools/gemm_time.py 128 huge 2>&1 | grep -E "^proj_ln|^fc2_ln" >> $O.gemm.txt
done
cat $O.gemm.txt
VPB_LN_FULLGRID=1 timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "layernorm" 2>&1 | tail -2
for f in 0 1 0 1; do
  VPB_LN_FULLGRID=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.bench.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.bench.json').read().strip().splitlines()[-1])
print('fullgrid=$f', round(r['value'],1), round(r['ms_per_step'],3), {k:round(v,4) for k,v in r['roofline']['ms_per_launch'].items() if 'ln' in k})"
done
