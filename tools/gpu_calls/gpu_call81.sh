#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_moe.py -x -q -m gpu > gpurun_out/r02_call81.moe.txt 2>&1; echo "moe rc=$?"; tail -25 gpurun_out/r02_call81.moe.txt
