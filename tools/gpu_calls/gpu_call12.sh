#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call12
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bwd_ops.py -x -q -m gpu -k "attention" > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
for d in 0 512 0 512; do VPB_ATT_DEBUG=$d timeout 120 python tools/att_time.py 512 64 >> $O.att.txt 2>&1; done
for d in 0 512; do VPB_ATT_DEBUG=$d timeout 120 python tools/att_time.py 512 32 >> $O.att.txt 2>&1; done
for d in 0 512; do VPB_ATT_DEBUG=$d timeout 120 python tools/att_time.py 128 64 >> $O.att.txt 2>&1; done
VPB_ATT_DEBUG=96 timeout 120 python tools/att_time.py 512 64 2>&1 | tail -3 >> $O.att.txt
tail -4 $O.tests.txt; cat $O.att.txt | cut -c1-600
