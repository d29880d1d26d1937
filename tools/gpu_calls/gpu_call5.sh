#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call5
for l in "" vitpose_b200/libvitpose_b200_s4.so "" vitpose_b200/libvitpose_b200_s4.so; do VPB_LIB=$l timeout 120 python tools/gemm_time.py 256 base 2>&1 | head -2 >> $O.gemm.txt; echo "lib=$l" >> $O.gemm.txt; done
VPB_LIB=vitpose_b200/libvitpose_b200_s4.so VPB_GEMM_DEBUG=1 timeout 120 python tools/gemm_time.py 256 base 2>&1 | grep "epi=8" | tail -2 >> $O.gemm.txt
VPB_LIB=vitpose_b200/libvitpose_b200_s4.so timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "layernorm" 2>&1 | tail -2 >> $O.gemm.txt
cat $O.gemm.txt
