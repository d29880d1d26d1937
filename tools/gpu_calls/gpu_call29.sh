#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call29
for f in 1 0; do VPB_PDL=0 VPB_TRAIN_FUSE=$f timeout 300 python tools/train_kernel_profile.py 64 5 > $O.kprof_fuse$f.txt 2>&1; grep -v Warn $O.kprof_fuse$f.txt | head -24 | cut -c1-130; done
