#!/bin/bash
# A/B: fused-LayerNorm short-K GEMMs (proj + LN, patch + LN) with 2 operand stages + 4-slot residual rings vs 3 + 3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call73
cp vitpose_b200/libvitpose_b200.so /tmp/lib_default.so
for rep in 1 2; do
  cp /tmp/lib_default.so vitpose_b200/libvitpose_b200.so
  python tools/gemm_time.py 256 base 2>&1 | grep -E "proj_ln|fc2_ln" | sed 's/^/default  /' | tee -a $O.gemm.txt
  cp vitpose_b200/libvpb_ss2.so vitpose_b200/libvitpose_b200.so
  python tools/gemm_time.py 256 base 2>&1 | grep -E "proj_ln|fc2_ln" | sed 's/^/stages2  /' | tee -a $O.gemm.txt
done
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "layernorm" 2>&1 | tail -2
for v in default stages2 default stages2; do
  if [ $v = default ]; then cp /tmp/lib_default.so vitpose_b200/libvitpose_b200.so; else cp vitpose_b200/libvpb_ss2.so vitpose_b200/libvitpose_b200.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.bench.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.bench.json').read().strip().splitlines()[-1])
print('$v', round(r['value'],1), round(r['ms_per_step'],3), 'proj', r['roofline']['ms_per_launch']['gemm_proj_ln'], 'patch', r['roofline']['ms_per_launch']['gemm_patch_ln'], 'e2e', round(r['e2e']['value'],1))" | tee -a $O.bench.txt
done
