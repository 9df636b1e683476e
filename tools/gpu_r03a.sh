#!/bin/bash
# round 2, GPU call 3A (1 GPU): cooperative Riesz kernel on 512-thread CTAs (128 registers, 8 / 4 pair terms in flight) -- parity, A/B, phases
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gd.py -m gpu -x -q > gpurun_out/r03a_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r03a_pytest.log
for nt in 1024 512 1024 512; do echo "== riesz_threads=$nt"; DZO_RIESZ_THREADS=$nt timeout 300 python tools/riesz_probe.py 4096 20 | tail -2; done 2>&1 | tee gpurun_out/r03a_riesz_ab.log
DZO_RIESZ_THREADS=512 timeout 300 python tools/riesz_phases.py 4096 20 > gpurun_out/r03a_riesz_phases.log 2>&1; cat gpurun_out/r03a_riesz_phases.log
