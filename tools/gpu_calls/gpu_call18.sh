#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call18
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "fold or layernorm" > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
grep -E "passed|failed|rc=|FAILED|Error|error" $O.tests.txt | tail -20
timeout 300 python tools/gemm_time.py 256 base > $O.gemm.txt 2>&1; tail -12 $O.gemm.txt
timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu -x > $O.model.txt 2>&1; echo "model rc=$?" >> $O.model.txt; tail -5 $O.model.txt
for f in 0 1 0 1; do
  VPB_LN_FOLD=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.bench_fold$f.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.bench_fold$f.json').read().strip().splitlines()[-1])
print('fold=$f', r['value'], r['ms_per_step'], r['e2e']['value'], r['clocks']['sm_mhz'])"
done
