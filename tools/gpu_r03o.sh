#!/bin/bash
# round 2, GPU call 3O (1 GPU): gradient items of two row blocks (shared staging, two rows per lane) -- parity, timing, phases
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gd.py -m gpu -x -q > gpurun_out/r03o_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r03o_pytest.log
for nt in 512 1024 512; do echo "== riesz_threads=$nt"; DZO_RIESZ_THREADS=$nt timeout 300 python tools/riesz_probe.py 4096 20 | tail -2; done 2>&1 | tee gpurun_out/r03o_riesz_ab.log
timeout 300 python tools/riesz_phases.py 4096 20 > gpurun_out/r03o_riesz_phases.log 2>&1; head -9 gpurun_out/r03o_riesz_phases.log
