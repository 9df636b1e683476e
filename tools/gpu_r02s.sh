#!/bin/bash
# round 2, GPU call S (1 GPU): batches of medium-n problems above the warp-search range on single-CTA search kernels; parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bfgs.py -m gpu -x -q -k "medium or large or warp or tuning" > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02s_pytest.log
timeout 600 python tools/medium_probe.py > gpurun_out/r02s_medium.log 2>&1; echo "probe rc=$?"
cut -c1-330 gpurun_out/r02s_medium.log
