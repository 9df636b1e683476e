#!/bin/bash
# round 2, GPU call K (1 GPU): warps per CTA of the batched kernel (1 / 2 / 4 / 8, same 16 warps per SM) A/B; Riesz phase log with the norms split out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=$PWD/dzoptimization.jl_b200/csrc/variants
for rep in 1 2; do
for v in default w2 w1 w8; do
  if [ $v == default ]; then unset DZOPT_B200_LIB; else export DZOPT_B200_LIB=$V/libdzopt_$v.so; fi
  echo "== $v"; timeout 300 python tools/batched_ab.py --configs l1p3 --repeat 1 2>/dev/null | cut -c1-140
done; done 2>&1 | tee gpurun_out/r02k_warps_ab.log
unset DZOPT_B200_LIB
DZO_RIESZ_PAIR=1 timeout 300 python tools/riesz_phases.py 4096 20 > gpurun_out/r02k_riesz_phases.log 2>&1; cat gpurun_out/r02k_riesz_phases.log
