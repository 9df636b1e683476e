#!/bin/bash
# round 2, GPU call 3P (1 GPU): problems per warp of the batched kernel chosen from the batch size -- parity, A/B at small batches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bfgs.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r03p_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r03p_pytest.log
for b in 1000000 500000 250000 125000 62500 20000; do
  for t in 32 0; do echo "== batch $b batched_tile=$t"; DZO_BATCHED_TILE=$t timeout 300 python tools/batched_ab.py --batch $b --configs l1p3 --repeat 1 2>/dev/null | cut -c1-110; done
done 2>&1 | tee gpurun_out/r03p_tile_ab.log
