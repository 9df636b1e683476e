#!/bin/bash
# round 2, GPU call 3V (1 GPU): fixed-tile and variable-tile instantiations of the batched kernel -- parity, timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bfgs.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r03v_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r03v_pytest.log
for b in 1000000 125000 1000000 125000; do echo "== batch $b"; timeout 300 python tools/batched_ab.py --batch $b --configs l1p3 --repeat 1 2>/dev/null | cut -c1-110; done
