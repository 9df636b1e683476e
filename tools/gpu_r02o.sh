#!/bin/bash
# round 2, GPU call O (1 GPU): grid-wide reductions through flagged lines instead of grid barriers -- parity, A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lbfgs.py tests/test_legacy_lbfgs.py tests/test_gpu_gd.py -m gpu -x -q > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02o_pytest.log
for ll in 0 1 0 1; do echo "== grid_ll=$ll"; DZO_GRID_LL=$ll timeout 120 python tools/lbfgs_probe.py | tail -3; done 2>&1 | tee gpurun_out/r02o_lbfgs_ab.log
for ll in 0 1; do echo "== n=2^22 grid_ll=$ll"; DZO_GRID_LL=$ll timeout 120 python tools/lbfgs_probe.py 4194304 | tail -2; done 2>&1 | tee -a gpurun_out/r02o_lbfgs_ab.log
for ll in 0 1; do echo "== legacy grid_ll=$ll"; DZO_GRID_LL=$ll timeout 120 python tools/legacy_probe.py | tail -2; done 2>&1 | tee -a gpurun_out/r02o_lbfgs_ab.log
