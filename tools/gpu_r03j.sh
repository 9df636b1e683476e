#!/bin/bash
# round 2, GPU call 3J (1 GPU): flag-word grid barrier in the Riesz k-step mode -- parity, A/B, phases
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gd.py -m gpu -x -q > gpurun_out/r03j_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r03j_pytest.log
for b in 0 1 0 1; do echo "== riesz_bar=$b"; DZO_RIESZ_BAR=$b timeout 300 python tools/riesz_probe.py 4096 20 | tail -2; done 2>&1 | tee gpurun_out/r03j_riesz_ab.log
DZO_RIESZ_BAR=1 timeout 300 python tools/riesz_phases.py 4096 20 > gpurun_out/r03j_riesz_phases.log 2>&1; head -9 gpurun_out/r03j_riesz_phases.log
