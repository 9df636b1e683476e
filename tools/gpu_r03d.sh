#!/bin/bash
# round 2, GPU call 3D (1 GPU): tag recycling of the flagged lines, reduction variants; all grid-kernel tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lbfgs.py tests/test_legacy_lbfgs.py -m gpu -x -q > gpurun_out/r03d_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r03d_pytest.log
timeout 120 python tools/lbfgs_probe.py | tail -2
