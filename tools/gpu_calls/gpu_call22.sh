#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call22
timeout 300 python bench.py --train --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.train.json 2>$O.err.txt; echo "rc=$?"; cut -c1-400 $O.train.json
VPB_COOP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O.launches_train.csv python bench.py --train --steps 2 --warmup 3 --no-extra --no-cpu-baseline > $O.ncu.log 2>&1; echo "ncu rc=$?"
wc -l $O.launches_train.csv
