#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_call60
python tools/attbwd_time.py 64 12 64 > $O.attbwd.txt 2>&1
VPB_ATTBWD_DEBUG=1 python tools/attbwd_time.py 64 12 64 2>&1 | tail -6 >> $O.attbwd.txt
python tools/attbwd_time.py 64 16 80 >> $O.attbwd.txt 2>&1
python tools/attbwd_time.py 64 12 32 >> $O.attbwd.txt 2>&1
cat $O.attbwd.txt
