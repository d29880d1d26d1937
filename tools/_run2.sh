cd /root/repo
timeout 600 python tools/check_determinism.py 2>&1 | tail -6 > gpurun_out/determinism.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_ddp.py > gpurun_out/check_ddp.log 2>&1
timeout 600 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_bwd_ops.py tests/test_training_ops.py -q -m gpu 2>&1 | tail -3 > gpurun_out/t_train.log
