#!/bin/bash
# A/B of the Riesz GD kernel variants inside one gpurun call (same box): lanes per row in the energy items, gradient batch.
V=dzoptimization.jl_b200/csrc/build/variants
echo "== default (esplit=2, grad batch 2)"; python tools/riesz_probe.py 4096 20 | tail -2
echo "== esplit=1"; DZO_RIESZ_ESPLIT=1 python tools/riesz_probe.py 4096 20 | tail -2
for gb in 3 4; do
  if [ -f $V/libdzopt_gb$gb.so ]; then echo "== grad batch $gb"; DZOPT_B200_LIB=$PWD/$V/libdzopt_gb$gb.so python tools/riesz_probe.py 4096 20 | tail -2; fi
done
