#!/bin/bash
# round-2 GPU call 1: attention epilogue warpgroup + GEMM pipeline flags, A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O.smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
for d in 0 512 256; do VPB_ATT_DEBUG=$d timeout 120 python tools/att_time.py 512 64 >> $O.att.txt 2>&1; done
for d in 0 512; do VPB_ATT_DEBUG=$d timeout 120 python tools/att_time.py 128 80 >> $O.att.txt 2>&1; done
for d in 0 512; do VPB_ATT_DEBUG=$d timeout 120 python tools/att_time.py 512 32 >> $O.att.txt 2>&1; done
VPB_ATT_DEBUG=96 timeout 120 python tools/att_time.py 512 64 >> $O.att.txt 2>&1
for f in 0 1 2 4 3 5 7; do VPB_GEMM_FLAGS=$f timeout 120 python tools/gemm_time.py 256 base >> $O.gemm.txt 2>&1; done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O.bench.json 2> $O.bench.err
VPB_GEMM_FLAGS=7 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O.bench_f7.json 2>> $O.bench.err
tail -3 $O.tests.txt; cat $O.att.txt $O.gemm.txt; cat $O.bench.json | head -c 600
