#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call36
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
grep -E "passed|failed|rc=|FAILED|Error|assert" $O.tests.txt | tail -8
for f in 1 0 1 0; do
  VPB_LN_BN192=$f timeout 300 python bench.py --workload S-classic-17 --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.S$f.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.S$f.json').read().strip().splitlines()[-1])
print('bn192=$f', round(r['value'],1), round(r['ms_per_step'],3), 'e2e', round(r['e2e']['value'],1), {k:round(v,4) for k,v in r['roofline']['ms_per_launch'].items() if k.startswith('gemm')})"
done
