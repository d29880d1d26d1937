cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "H-classic" 2>&1 | tail -5 > gpurun_out/t_H.log
timeout 300 python bench.py --workload H-classic-133 --crops 64 --no-cpu-baseline --steps 5 > gpurun_out/bench_H.json 2> gpurun_out/bench_H.err
