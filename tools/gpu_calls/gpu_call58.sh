#!/bin/bash
# attention backward for head_dim 32 / 80, training step for ViTPose-S / -H
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bwd_ops.py -x -q -m gpu -k attention_bwd > gpurun_out/r02_call58.attbwd.txt 2>&1
echo "attbwd rc=$?"; tail -30 gpurun_out/r02_call58.attbwd.txt
timeout 900 python -m pytest tests/test_gpu_train_step.py -q -m gpu > gpurun_out/r02_call58.train.txt 2>&1
echo "train rc=$?"; tail -40 gpurun_out/r02_call58.train.txt
