#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/profile_overhead.py > gpurun_out/r02_call87.txt 2>gpurun_out/r02_call87.err; tail -5 gpurun_out/r02_call87.txt; tail -3 gpurun_out/r02_call87.err
