#!/bin/bash
# round 2, GPU call 3K (1 GPU): degenerate / awkward inputs on the warp-search sizes; new Riesz variant test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bfgs.py tests/test_gpu_gd.py -m gpu -x -q -k "degenerate or awkward or flag_word" > gpurun_out/r03k_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r03k_pytest.log
