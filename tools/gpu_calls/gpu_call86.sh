#!/bin/bash
# final validation of the round: smoke, full GPU suite, default bench (with sub-records), reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call86
python -c "import __graft_entry__ as g; g.smoke()" > $O.smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O.smoke.txt
timeout 1200 python -m pytest tests -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?"; tail -3 $O.tests.txt
python bench.py > $O.bench_default.json 2> $O.bench_default.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$O.bench_default.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), d['clocks'])
print('frac', d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['transformer_gemms_tflops'])
for n,r in d['configs'].items(): print(n, round(r['value'],1), round(r['ms_per_step'],3), 'e2e', round(r['e2e']['value'],1))
print(d['cpu_baseline'])"
