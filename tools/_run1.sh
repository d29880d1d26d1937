cd /root/repo
timeout 600 python tools/check_determinism.py > gpurun_out/determinism.log 2>&1
