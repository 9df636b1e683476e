#!/bin/bash
# round 2, GPU call L (1 GPU): warp-per-problem O(n) stage for medium n -- parity, timing; Riesz norms pass
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bfgs.py tests/test_gpu_golden.py tests/test_gpu_gd.py -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02l_pytest.log
timeout 300 python tools/medium_probe.py > gpurun_out/r02l_medium.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r02l_medium.log | cut -c1-250
DZO_RIESZ_PAIR=1 timeout 300 python tools/riesz_phases.py 4096 20 > gpurun_out/r02l_riesz_phases.log 2>&1; head -8 gpurun_out/r02l_riesz_phases.log
timeout 300 python tools/riesz_probe.py 4096 20 | tail -2
