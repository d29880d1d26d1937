#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call30
timeout 900 python -m pytest tests/test_gpu_bwd_ops.py tests/test_gpu_train_step.py tests/test_gpu_ops.py -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
grep -E "passed|failed|rc=|FAILED|Error" $O.tests.txt | tail -12
for f in 1 0 1 0; do
  VPB_TRAIN_FUSE=$f timeout 300 python bench.py --train --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.train_fuse$f.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.train_fuse$f.json').read().strip().splitlines()[-1])
print('fuse=$f', r['value'], r['ms_per_step'], r['e2e']['value'], r['gpu_launches'], r['final_loss'])"
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.bench.json 2>>$O.err.txt
python -c "
import json
r=json.loads(open('$O.bench.json').read().strip().splitlines()[-1])
print('infer', r['value'], r['ms_per_step'], r['e2e']['value'], {k:round(v,4) for k,v in r['roofline']['ms_per_launch'].items() if k.startswith('gemm')})"
