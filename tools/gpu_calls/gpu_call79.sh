#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call79
for c in 256 32 48 64 96 128 256; do
  timeout 300 python bench.py --crops $c --steps 20 --warmup 5 --no-extra --no-cpu-baseline > $O.bench.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.bench.json').read().strip().splitlines()[-1])
pl=r['roofline']['ms_per_launch']
s=256/$c
print('crops=$c', round(r['value'],1), round(r['ms_per_step'],3), 'per-256-equivalent ms: qkv %.3f att %.3f proj %.3f fc1 %.3f fc2 %.3f deconv %.3f' % (pl['gemm_qkv']*s, pl['attention']*s, pl['gemm_proj_ln']*s, pl['gemm_fc1']*s, pl['gemm_fc2_ln']*s, pl['deconv']*s), r['clocks']['sm_mhz'])" | tee -a $O.bench.txt
done
