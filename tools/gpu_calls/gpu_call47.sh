#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call47
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
grep -E "passed|failed|rc=|FAILED|Error|assert" $O.tests.txt | tail -6
for f in 0 1 0 1; do
  VPB_LN_FULLGRID=$f timeout 400 python bench.py --workload H-classic-133 --crops 256 --steps 3 --warmup 3 --no-extra --no-cpu-baseline > $O.H$f.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.H$f.json').read().strip().splitlines()[-1])
print('H fullgrid=$f', round(r['value'],1), round(r['ms_per_step'],2), 'e2e', round(r['e2e']['value'],1), {k:round(v,4) for k,v in r['roofline']['ms_per_launch'].items() if 'ln' in k})"
done
