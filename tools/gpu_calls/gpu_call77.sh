#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call77
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_bwd_ops.py -x -q -m gpu > $O.test.txt 2>&1; echo "test rc=$?"; tail -4 $O.test.txt
for v in 0 1 0 1; do
  VPB_DECONV_WIDE=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.bench.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.bench.json').read().strip().splitlines()[-1])
print('deconv_wide=$v', round(r['value'],1), round(r['ms_per_step'],3), 'deconv', r['roofline']['ms_per_launch']['deconv'], 'e2e', round(r['e2e']['value'],1), r['clocks']['sm_mhz'])" | tee -a $O.bench.txt
done
