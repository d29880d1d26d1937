#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call2
tools/bin/tmem_bw > $O.tmem.txt 2>&1
for d in 0 1024 1; do VPB_ATT_DEBUG=$d timeout 120 python tools/att_time.py 512 64 >> $O.att.txt 2>&1; done
VPB_GEMM_DEBUG=1 timeout 120 python tools/gemm_time.py 256 base > $O.gemmdbg.txt 2>&1
timeout 120 python tools/gemm_time.py 256 base > $O.gemm.txt 2>&1
cat $O.tmem.txt $O.att.txt $O.gemm.txt; grep "epi=8" $O.gemmdbg.txt | tail -4; grep "epi=6" $O.gemmdbg.txt | tail -3; grep "epi=1" $O.gemmdbg.txt | tail -3
