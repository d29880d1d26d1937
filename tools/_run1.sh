cd /root/repo
timeout 300 python tools/train_profile.py 64 2 > gpurun_out/train_prof.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches2.csv python tools/train_profile.py 64 2 > gpurun_out/ncu_train.log 2>&1
