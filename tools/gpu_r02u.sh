#!/bin/bash
# round 2, GPU call U (1 GPU): why does bench.py's live L-BFGS leg see 0.20 ms per step! when the probe sees 0.13?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== probe seed 1"; timeout 120 python tools/lbfgs_probe.py | tail -2
echo "== probe seed 9"; DZO_SEED=9 timeout 120 python tools/lbfgs_probe.py | tail -2
echo "== probe seed 9 set_stream"; DZO_SEED=9 DZO_SETSTREAM=1 timeout 120 python tools/lbfgs_probe.py | tail -2
echo "== probe seed 9 stage 0"; DZO_SEED=9 DZO_GRID_STAGE=0 timeout 120 python tools/lbfgs_probe.py | tail -2
echo "== probe seed 9 ll 0 stage 0"; DZO_SEED=9 DZO_GRID_STAGE=0 DZO_GRID_LL=0 timeout 120 python tools/lbfgs_probe.py | tail -2
echo "== probe seed 9 ll 0 stage 1"; DZO_SEED=9 DZO_GRID_STAGE=1 DZO_GRID_LL=0 timeout 120 python tools/lbfgs_probe.py | tail -2
