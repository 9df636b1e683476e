#!/bin/bash
# round 2, GPU call Y (1 GPU): compile-time variants of the flagged-line reductions side by side on ONE box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=$PWD/dzoptimization.jl_b200/csrc/variants
for rep in 1; do
for lib in default early smem earlysmem rstate; do
  if [ $lib == default ]; then unset DZOPT_B200_LIB; else export DZOPT_B200_LIB=$V/libdzopt_$lib.so; fi
  for cfg in "1 1" "0 1"; do set -- $cfg; echo "== $lib lbfgs grid_ll=$1 grid_stage=$2"; DZO_GRID_LL=$1 DZO_GRID_STAGE=$2 timeout 120 python tools/lbfgs_probe.py | tail -2 | head -1; done
  echo "== $lib lbfgs 2^22 ll=1"; DZO_GRID_LL=1 timeout 120 python tools/lbfgs_probe.py 4194304 | tail -2 | head -1
  echo "== $lib legacy ll=1"; DZO_GRID_LL=1 timeout 120 python tools/legacy_probe.py | tail -1
done; done 2>&1 | tee gpurun_out/r02y_variants.log
