#!/bin/bash
# round 2, GPU call P (1 GPU): live L-BFGS -- next pass's vectors staged into shared memory behind the reduction; parity + A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lbfgs.py -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02p_pytest.log
for cfg in "0 0" "1 0" "1 1" "0 1" "1 1"; do set -- $cfg; echo "== grid_ll=$1 grid_stage=$2"; DZO_GRID_LL=$1 DZO_GRID_STAGE=$2 timeout 120 python tools/lbfgs_probe.py | tail -3; done 2>&1 | tee gpurun_out/r02p_lbfgs_ab.log
