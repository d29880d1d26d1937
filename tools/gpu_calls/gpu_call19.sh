#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call19
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "fold" > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
grep -E "passed|failed|rc=|FAILED" $O.tests.txt | tail -5
timeout 300 python tools/gemm_time.py 256 base > $O.gemm.txt 2>&1; tail -8 $O.gemm.txt
for f in 0 1 0 1; do
  VPB_LN_FOLD=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.bench_fold$f.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.bench_fold$f.json').read().strip().splitlines()[-1])
print('fold=$f', r['value'], r['ms_per_step'], r['e2e']['value'], r['clocks']['sm_mhz'], {k:round(v,4) for k,v in r['roofline']['ms_per_launch'].items() if k.startswith('gemm')})"
done
