#!/bin/bash
# round 2, GPU call 3F (1 GPU): legacy grid-wide L-BFGS with staged passes -- parity, A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_legacy_lbfgs.py tests/test_lbfgs.py tests/test_gpu_gd.py -m gpu -x -q > gpurun_out/r03f_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r03f_pytest.log
for st in 0 1 0 1; do echo "== legacy grid_stage=$st"; DZO_GRID_STAGE=$st timeout 120 python tools/legacy_probe.py | tail -1; done 2>&1 | tee gpurun_out/r03f_legacy_ab.log
echo "== legacy n=2^22"; timeout 120 python tools/legacy_probe.py 4194304 | tail -1
