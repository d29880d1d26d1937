#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call17
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bwd_ops.py tests/test_gpu_model.py -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
for l in "" vitpose_b200/libvitpose_b200_cv1.so "" vitpose_b200/libvitpose_b200_cv1.so; do VPB_LIB=$l timeout 200 python tools/bench_ops.py --crops 256 2>&1 | grep -E "deconv|final1x1" >> $O.ops.txt; echo "lib=$l" >> $O.ops.txt; done
cat $O.ops.txt; grep -E "passed|failed|rc=|FAILED|Error" $O.tests.txt | tail
