#!/bin/bash
# round 2, GPU call X (1 GPU): staged fetch issued after the CTA's line stores
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lbfgs.py -m gpu -x -q 2>&1 | tail -2
for cfg in "1 1" "0 1" "1 1" "0 1"; do set -- $cfg; echo "== lbfgs grid_ll=$1 grid_stage=$2"; DZO_GRID_LL=$1 DZO_GRID_STAGE=$2 timeout 120 python tools/lbfgs_probe.py | tail -2; done 2>&1 | tee gpurun_out/r02x_lbfgs.log
for ll in 1 0; do echo "== n=2^22 ll=$ll"; DZO_GRID_LL=$ll timeout 120 python tools/lbfgs_probe.py 4194304 | tail -2 | head -1; done 2>&1 | tee -a gpurun_out/r02x_lbfgs.log
