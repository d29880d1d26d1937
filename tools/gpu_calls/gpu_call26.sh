#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call26
for f in 1 0; do VPB_TRAIN_FUSE=$f timeout 300 python tools/train_cpu_probe.py 64 >> $O.probe.txt 2>&1; VPB_TRAIN_FUSE=$f timeout 300 python tools/train_cpu_probe.py 128 >> $O.probe.txt 2>&1; done
cat $O.probe.txt
