#!/bin/bash
# round 2, GPU call W (1 GPU): poll back-off of the flagged-line reductions
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for bo in 0 32 100 300 1000; do
  echo "== lbfgs ll=1 stage=1 backoff=$bo"; DZO_GRID_BACKOFF=$bo timeout 120 python tools/lbfgs_probe.py | tail -2 | head -1
  echo "== legacy ll=1 backoff=$bo"; DZO_GRID_BACKOFF=$bo timeout 120 python tools/legacy_probe.py | tail -1
done 2>&1 | tee gpurun_out/r02w_backoff.log
echo "== lbfgs ll=1 stage=0"; DZO_GRID_STAGE=0 timeout 120 python tools/lbfgs_probe.py | tail -2 | head -1
