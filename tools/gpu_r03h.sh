#!/bin/bash
# round 2, GPU call 3H (1 GPU): cache operators of the update sweep's H loads / stores
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=$PWD/dzoptimization.jl_b200/csrc/variants
for rep in 1 2; do
for lib in default st1 st2 st3 ld1 ld3 ld2st1; do
  if [ $lib == default ]; then unset DZOPT_B200_LIB; else export DZOPT_B200_LIB=$V/libdzopt_$lib.so; fi
  timeout 120 python tools/sweep_probe2.py
done; done 2>&1 | tee gpurun_out/r03h_cacheops.jsonl
