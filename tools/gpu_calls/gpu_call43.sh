#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call43
export VPB_COOP=0
CMD="python bench.py --steps 2 --warmup 3 --crops 64 --no-extra --no-cpu-baseline"
$CMD > $O.plain_b64.json 2> $O.plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O.launches_b64.csv $CMD > $O.ncu1.log 2>&1
CMD2="python bench.py --steps 1 --warmup 3 --crops 256 --no-extra --no-cpu-baseline"
$CMD2 > $O.plain_b256.json 2>> $O.plain.err && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_kernel -s 200 -c 6 -o $O.prof_gemms $CMD2 > $O.ncu2.log 2>&1
ls -la gpurun_out/ | grep r02_call43
tail -3 $O.ncu1.log $O.ncu2.log
