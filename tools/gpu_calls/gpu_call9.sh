#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call9
timeout 1500 python -m pytest tests -q -m gpu -s > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
for f in 1 0 1 0; do VPB_SIMPLE_FUSED=$f timeout 300 python bench.py --workload L-simple-17 --crops 512 --steps 5 --warmup 3 --no-cpu-baseline --no-extra > $O.L_fused$f.json 2>> $O.bench.err; python -c "
import json
d=json.loads(open('$O.L_fused$f.json').read().strip().splitlines()[-1])
print('L fused=$f', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['roofline']['ms_per_launch'])
"; done
grep -E "passed|failed|rc=|peaked|well-posed|FAILED|Error" $O.tests.txt | tail -40
