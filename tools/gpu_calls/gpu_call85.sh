#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call85
timeout 600 python -m pytest tests/test_gpu_bwd_ops.py tests/test_gpu_train_step.py -x -q -m gpu > $O.test.txt 2>&1; echo "test rc=$?"; tail -3 $O.test.txt
VPB_PDL=0 timeout 300 python tools/train_kernel_profile.py 64 5 > $O.train_kernels.txt 2>&1; grep -E "kernels busy|layernorm_bwd" $O.train_kernels.txt
