#!/bin/bash
# round 2, GPU call V (1 GPU): grid-wide L-BFGS with gpu-scope relaxed line loads / stores; with and without the profile code compiled in
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=$PWD/dzoptimization.jl_b200/csrc/variants
for lib in default noprof default noprof; do
  if [ $lib == default ]; then unset DZOPT_B200_LIB; else export DZOPT_B200_LIB=$V/libdzopt_$lib.so; fi
  for cfg in "1 1" "0 1"; do set -- $cfg; echo "== $lib grid_ll=$1 grid_stage=$2"; DZO_GRID_LL=$1 DZO_GRID_STAGE=$2 timeout 120 python tools/lbfgs_probe.py | tail -2; done
  echo "== $lib legacy ll=1"; DZO_GRID_LL=1 timeout 120 python tools/legacy_probe.py | tail -1
  echo "== $lib legacy ll=0"; DZO_GRID_LL=0 timeout 120 python tools/legacy_probe.py | tail -1
done 2>&1 | tee gpurun_out/r02v_lbfgs.log
unset DZOPT_B200_LIB
timeout 900 python -m pytest tests/test_lbfgs.py tests/test_legacy_lbfgs.py -m gpu -x -q 2>&1 | tail -2
