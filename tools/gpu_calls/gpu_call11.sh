#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call11
for d in 0 2048 0 2048 1 3 7 8 9 15; do VPB_ATT_DEBUG=$d timeout 120 python tools/att_time.py 512 64 >> $O.att.txt 2>&1; done
cat $O.att.txt
