#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call13
for d in 512 2560 512 2560; do VPB_ATT_DEBUG=$d timeout 120 python tools/att_time.py 512 64 >> $O.att.txt 2>&1; done
VPB_ATT_DEBUG=2656 timeout 120 python tools/att_time.py 512 64 2>&1 | tail -3 >> $O.att.txt
cat $O.att.txt | cut -c1-500
