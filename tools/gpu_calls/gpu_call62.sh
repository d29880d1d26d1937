#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_call62
timeout 600 python -m pytest tests/test_gpu_bwd_ops.py -x -q -m gpu -k attention_bwd > $O.test.txt 2>&1; echo "test rc=$?"; tail -15 $O.test.txt
python tools/attbwd_time.py 64 12 64 > $O.attbwd.txt 2>&1
VPB_ATTBWD_PREFETCH=0 python tools/attbwd_time.py 64 12 64 >> $O.attbwd.txt 2>&1
VPB_ATTBWD_PREFETCH=296 python tools/attbwd_time.py 64 12 64 >> $O.attbwd.txt 2>&1
VPB_ATTBWD_DEBUG=500 python tools/attbwd_time.py 64 12 64 2>&1 | tail -3 >> $O.attbwd.txt
python tools/attbwd_time.py 64 16 80 >> $O.attbwd.txt 2>&1
python tools/attbwd_time.py 64 12 32 >> $O.attbwd.txt 2>&1
cat $O.attbwd.txt
