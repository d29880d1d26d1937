#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call54
timeout 900 python -m pytest tests/test_moe.py tests/test_gpu_model.py -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?" >> $O.tests.txt
grep -E "passed|failed|rc=|FAILED|Error|assert" $O.tests.txt | tail -12
