#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call32
for w in H-classic-133 L-simple-17 S-classic-17; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-extra --no-cpu-baseline > $O.$w.json 2>$O.err.txt
  python -c "
import json
r=json.loads(open('$O.$w.json').read().strip().splitlines()[-1])
rf=r['roofline']
print('$w', r['config'].get('crops_per_gpu'), round(r['value'],1), round(r['ms_per_step'],2), 'e2e', round(r['e2e']['value'],1), 'frac', round(rf['frac'],3), rf.get('transformer_gemms_tflops'))
print('   ', {k:round(v,4) for k,v in rf['ms_per_launch'].items()})
print('   ', {k:round(v,4) for k,v in rf['step_share_by_kernel'].items()})"
done
