#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call88
python bench.py > $O.bench_default.json 2> $O.bench_default.err; echo "bench rc=$?"; tail -3 $O.bench_default.err
python -c "
import json
d=json.loads(open('$O.bench_default.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), d['clocks'])
r=d['roofline']
print('frac', r['frac'], r['achieved'], r['transformer_gemms_tflops'], 'with events', r['ms_per_step_with_events'], 'shares sum', sum(r['step_share_by_kernel'].values()))
for n,c in d['configs'].items(): print(n, round(c['value'],1), round(c['ms_per_step'],3), 'e2e', round(c['e2e']['value'],1))"
