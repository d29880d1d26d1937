#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call3
for f in 4 12 20; do VPB_GEMM_FLAGS=$f timeout 120 python tools/gemm_time.py 256 base >> $O.gemm.txt 2>&1; done
cat $O.gemm.txt
