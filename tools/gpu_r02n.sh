#!/bin/bash
# round 2, GPU call N (1 GPU): sensitivity of the batched kernel to resident warps per SM (shared-memory padding: 16 -> 12 -> 8 warps)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for pad in 0 20000 0 20000 60000; do
  echo "== pad $pad"; DZO_HYBRID_SMEM_PAD=$pad timeout 300 python tools/batched_ab.py --configs l1p3 --repeat 1 2>/dev/null | cut -c1-140
done 2>&1 | tee gpurun_out/r02n_occupancy.log
