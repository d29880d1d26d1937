#!/bin/bash
# full GPU suite after the ViTPose+ training / attention-backward changes; training kernel table; training bench
mkdir -p gpurun_out
O=gpurun_out/r02_call59
timeout 1200 python -m pytest tests -x -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?"; tail -4 $O.tests.txt
VPB_PDL=0 timeout 300 python tools/train_kernel_profile.py 64 5 > $O.train_kernels.txt 2>&1; head -16 $O.train_kernels.txt
for i in 1 2; do
timeout 300 python bench.py --train --steps 10 --warmup 3 --no-cpu-baseline > $O.train$i.json 2>$O.err.txt
python -c "
import json
r=json.loads(open('$O.train$i.json').read().strip().splitlines()[-1])
print('train', round(r['value'],1), round(r['ms_per_step'],3))" || tail -3 $O.err.txt
done
