#!/bin/bash
# round 2, GPU call Q (1 GPU): cycle split of the grid-wide L-BFGS kernel's leader CTA
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in "1 1" "1 0" "0 1"; do set -- $cfg; echo "== grid_ll=$1 grid_stage=$2"; DZO_GRID_PROFILE=1 DZO_GRID_LL=$1 DZO_GRID_STAGE=$2 timeout 120 python tools/lbfgs_probe.py 1048576 10 50 | tail -6; done 2>&1 | tee gpurun_out/r02q_lbfgs_profile.log
