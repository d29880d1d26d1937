cd /root/repo
timeout 900 python -m pytest tests/test_gpu_bwd_ops.py -q -m gpu 2>&1 | tail -40 > gpurun_out/t_bwd.log
