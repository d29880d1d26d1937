#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call6
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
for l in "" vitpose_b200/libvitpose_b200_nodec.so vitpose_b200/libvitpose_b200_sb4.so "" vitpose_b200/libvitpose_b200_nodec.so; do VPB_LIB=$l timeout 120 python tools/gemm_time.py 256 base 2>&1 >> $O.gemm.txt; echo "lib=$l" >> $O.gemm.txt; done
VPB_GEMM_DEBUG=1 timeout 120 python tools/gemm_time.py 256 base > $O.gemmdbg.txt 2>&1
tail -3 $O.tests.txt; cat $O.gemm.txt; grep "epi=1" $O.gemmdbg.txt | tail -2; grep "epi=8" $O.gemmdbg.txt | tail -2;  grep "epi=6" $O.gemmdbg.txt | tail -2; grep "epi=0" $O.gemmdbg.txt | tail -2
