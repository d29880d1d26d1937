#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call10
timeout 1500 python -m pytest tests -q -m gpu -s > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
timeout 300 python bench.py --workload L-simple-17 --crops 512 --steps 5 --warmup 3 --no-cpu-baseline --no-extra > $O.L.json 2>> $O.bench.err
for f in 1 0 1 0; do VPB_GEMM_WIDE=$f timeout 300 python bench.py --workload S-classic-17 --crops 256 --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O.S_wide$f.json 2>> $O.bench.err; done
python - <<PY
import json
for f in ['L','S_wide1','S_wide0']:
    d=json.loads(open('$O.'+f+'.json').read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['roofline']['ms_per_launch'])
PY
grep -E "passed|failed|rc=|FAILED|Error" $O.tests.txt | tail -20
