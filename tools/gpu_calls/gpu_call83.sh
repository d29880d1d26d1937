#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call83
timeout 600 python -m pytest tests/test_gpu_bwd_ops.py tests/test_gpu_train_step.py -x -q -m gpu > $O.test.txt 2>&1; echo "test rc=$?"; tail -3 $O.test.txt
VPB_PDL=0 timeout 300 python tools/train_kernel_profile.py 64 5 > $O.train_kernels.txt 2>&1; grep -E "kernels busy|deconv_gather|deconv_phase" $O.train_kernels.txt
for i in 1 2; do
timeout 300 python bench.py --train --steps 10 --warmup 3 --no-cpu-baseline > $O.train$i.json 2>$O.err.txt
python -c "
import json
r=json.loads(open('$O.train$i.json').read().strip().splitlines()[-1])
print('train', round(r['value'],1), round(r['ms_per_step'],3), 'e2e', round(r['e2e']['ms_per_step'],3))" || tail -3 $O.err.txt
done
