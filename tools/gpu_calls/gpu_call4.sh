#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call4
VPB_GEMM_HI=1 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm or layernorm" > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
for h in 0 1 0 1; do VPB_GEMM_HI=$h timeout 120 python tools/gemm_time.py 256 base >> $O.gemm.txt 2>&1; echo "HI=$h" >> $O.gemm.txt; done
VPB_GEMM_HI=1 VPB_GEMM_DEBUG=1 timeout 120 python tools/gemm_time.py 256 base > $O.gemmdbg.txt 2>&1
tail -3 $O.tests.txt; cat $O.gemm.txt; grep "epi=1" $O.gemmdbg.txt | tail -2; grep "epi=8" $O.gemmdbg.txt | tail -2
