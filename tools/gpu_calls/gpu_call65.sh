#!/bin/bash
# ncu evidence for the reworked attention backward (train step), launch list of the train step, full default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call65
python bench.py > $O.bench_default.json 2> $O.bench_default.err; echo "bench rc=$?"; tail -c 600 $O.bench_default.json
export VPB_COOP=0
CMD="python bench.py --train --steps 1 --warmup 3 --no-extra --no-cpu-baseline"
$CMD > $O.plain.json 2> $O.plain.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_bwd" -s 12 -c 2 -o $O.prof_attbwd $CMD > $O.ncu.log 2>&1
tail -n 2 $O.ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O.launches_train.csv $CMD > $O.ncu2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ | grep r02_call65
