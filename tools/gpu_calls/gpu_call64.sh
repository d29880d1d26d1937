#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_call64
timeout 600 python -m pytest tests/test_gpu_bwd_ops.py -x -q -m gpu -k attention_bwd > $O.test.txt 2>&1; echo "test rc=$?"; tail -15 $O.test.txt
python tools/attbwd_time.py 64 12 64 > $O.attbwd.txt 2>&1
VPB_ATTBWD_DEBUG=500 python tools/attbwd_time.py 64 12 64 2>&1 | tail -3 >> $O.attbwd.txt
python tools/attbwd_time.py 64 16 80 >> $O.attbwd.txt 2>&1
python tools/attbwd_time.py 64 12 32 >> $O.attbwd.txt 2>&1
cat $O.attbwd.txt
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_moe.py -x -q -m gpu > $O.train.txt 2>&1; echo "train rc=$?"; tail -5 $O.train.txt
for i in 1 2; do
timeout 300 python bench.py --train --steps 10 --warmup 3 --no-cpu-baseline > $O.train$i.json 2>$O.err.txt
python -c "
import json
r=json.loads(open('$O.train$i.json').read().strip().splitlines()[-1])
print('train', round(r['value'],1), round(r['ms_per_step'],3))" || tail -3 $O.err.txt
done
