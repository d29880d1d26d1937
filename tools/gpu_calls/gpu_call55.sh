#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call55
timeout 1500 python -m pytest tests -q -m gpu -x > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
tail -4 $O.tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O.smoke.txt 2>&1; tail -2 $O.smoke.txt
timeout 900 python bench.py > $O.bench.json 2>$O.bench.err; echo "bench rc=$?"
python -c "
import json
r=json.loads(open('$O.bench.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], r['e2e'], r['clocks'], r['roofline'], r['cpu_baseline'])
for k,v in (r.get('configs') or {}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e'))
"
