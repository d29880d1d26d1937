#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02_call68
python -c "import __graft_entry__ as g; g.smoke()" > $O.smoke.txt 2>&1; echo "smoke rc=$?"; tail -5 $O.smoke.txt
timeout 1200 python -m pytest tests -q -m gpu > $O.tests.txt 2>&1; echo "tests rc=$?"; tail -4 $O.tests.txt
