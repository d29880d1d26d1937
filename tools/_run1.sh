cd /root/repo
timeout 600 python -m pytest tests/test_moe.py -q -m gpu 2>&1 | tail -25 > gpurun_out/t_moe.log
