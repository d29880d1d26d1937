cd /root/repo
timeout 900 python -m pytest tests/test_gpu_bwd_ops.py tests/test_gpu_train_step.py -q -m gpu 2>&1 | tail -12 > gpurun_out/t_train.log
timeout 300 python tools/train_profile.py 64 4 2>&1 | tail -2 > gpurun_out/train_prof.log
timeout 600 python bench.py --train --steps 5 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
