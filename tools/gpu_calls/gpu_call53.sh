#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call53
timeout 900 python -m pytest tests/test_gpu_bwd_ops.py tests/test_gpu_train_step.py -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
grep -E "passed|failed|rc=|FAILED|Error|assert" $O.tests.txt | tail -6
VPB_PDL=0 timeout 300 python tools/train_kernel_profile.py 64 5 2>&1 | grep -E "attention_bwd|kernels busy|colsum"
for i in 1 2; do
  timeout 300 python bench.py --train --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.train.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.train.json').read().strip().splitlines()[-1])
print('train', round(r['value'],1), round(r['ms_per_step'],3), 'e2e', round(r['e2e']['value'],1))"
done
