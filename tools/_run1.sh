cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_train_step.py -q -m gpu 2>&1 | tail -15 > gpurun_out/t_train.log
