#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call52
export VPB_COOP=0
CMD="python bench.py --train --steps 1 --warmup 3 --no-extra --no-cpu-baseline"
$CMD > $O.plain.json 2> $O.plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tn_kernel<256, 1[012]|attention_bwd|layernorm_bwd|cast_colsum" -s 150 -c 10 -o $O.prof_train $CMD > $O.ncu.log 2>&1
ls -la gpurun_out/ | grep r02_call52
tail -n 3 $O.ncu.log
