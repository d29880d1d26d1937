#!/bin/bash
# ViTPose+ training step: parity vs the oracle; then the whole training test file
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_moe.py -x -q -m gpu > gpurun_out/r02_call57.moe.txt 2>&1
echo "moe rc=$?"; tail -30 gpurun_out/r02_call57.moe.txt
timeout 600 python -m pytest tests/test_gpu_train_step.py -x -q -m gpu > gpurun_out/r02_call57.train.txt 2>&1
echo "train rc=$?"; tail -5 gpurun_out/r02_call57.train.txt
