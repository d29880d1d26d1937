#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call20
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "gelu or fold" > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
grep -E "passed|failed|rc=|FAILED|assert" $O.tests.txt | tail -8
for l in "" vitpose_b200/libvitpose_b200_as.so "" vitpose_b200/libvitpose_b200_as.so; do
  echo "lib=$l" >> $O.gemm.txt
  VPB_LIB=$l timeout 300 python tools/gemm_time.py 256 base 2>&1 | grep -E "^fc1 |^qkv " >> $O.gemm.txt
done
cat $O.gemm.txt
for l in "" vitpose_b200/libvitpose_b200_as.so "" vitpose_b200/libvitpose_b200_as.so; do
  VPB_LIB=$l VPB_LN_FOLD=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O.bench.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.bench.json').read().strip().splitlines()[-1])
print('lib=$l', r['value'], r['ms_per_step'], r['e2e']['value'], r['clocks']['sm_mhz'], {k:round(v,4) for k,v in r['roofline']['ms_per_launch'].items() if k.startswith('gemm')})"
done
