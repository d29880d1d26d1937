#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call7
timeout 1500 python -m pytest tests -x -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
VPB_PDL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O.bench_nopdl.json 2> $O.bench.err
VPB_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O.bench_pdl.json 2>> $O.bench.err
VPB_PDL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O.bench_nopdl2.json 2>> $O.bench.err
VPB_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O.bench_pdl2.json 2>> $O.bench.err
tail -5 $O.tests.txt
for f in nopdl pdl nopdl2 pdl2; do python -c "
import json,sys
d=json.loads(open('$O.bench_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'], d['roofline']['ms_per_launch'])
"; done
