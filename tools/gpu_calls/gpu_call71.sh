#!/bin/bash
# ncu --set full of the dominant kernel (fc1 + GELU) at the bench's full size: live DRAM traffic for roofline.traffic
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call71
export VPB_COOP=0
CMD="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline"
$CMD > $O.plain.json 2> $O.plain.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_kernel -s 150 -c 8 -o $O.prof_fc1 $CMD > $O.ncu.log 2>&1
tail -n 2 $O.ncu.log
ls -la $O.*
