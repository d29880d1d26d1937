cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_bwd -c 2 -o gpurun_out/prof_attnbwd -f python tools/prof_gemm.py attn_bwd > gpurun_out/ncu_attnbwd.log 2>&1
