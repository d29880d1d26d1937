#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call27
for n in 2 16; do timeout 300 python tools/train_cpu_probe.py $n >> $O.probe.txt 2>&1; done
cat $O.probe.txt
timeout 300 python - > $O.cprof.txt 2>&1 <<'EOF'
import cProfile, pstats, sys, os, torch
sys.argv = ['x', '2']
sys.path.insert(0, 'tools')
import importlib.util
spec = importlib.util.spec_from_file_location('probe', 'tools/train_cpu_probe.py')
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): m.step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
EOF
head -90 $O.cprof.txt | cut -c1-160
