#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call16
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bwd_ops.py -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "small_config or golden" >> $O.tests.txt 2>&1; echo "tests2 rc=$?" >> $O.tests.txt
for f in 1 0 1 0; do VPB_GEMM_PAIR128=$f timeout 300 python bench.py --workload S-classic-17 --crops 256 --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O.S_pair$f.json 2>> $O.bench.err; python -c "
import json
d=json.loads(open('$O.S_pair$f.json').read().strip().splitlines()[-1])
print('S pair128=$f', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['roofline']['ms_per_launch'])
"; done
VPB_LOG_ON_DEVICE=0 timeout 300 python bench.py --train --steps 20 --warmup 3 > $O.train_sync.json 2>> $O.bench.err
VPB_LOG_ON_DEVICE=1 timeout 300 python bench.py --train --steps 20 --warmup 3 > $O.train_nosync.json 2>> $O.bench.err
for f in train_sync train_nosync; do python -c "
import json
d=json.loads(open('$O.$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'])
"; done
grep -E "passed|failed|rc=|FAILED|Error" $O.tests.txt | tail
