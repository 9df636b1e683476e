#!/bin/bash
# round 2, GPU call C (1 GPU): batched kernel with asynchronous staging -- parity subset, A/B, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bfgs.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02d_pytest.log
timeout 600 python tools/batched_ab.py --configs v2l0p2,v0l1p3s0,v0l1p3s1,v0l1p2s1,v0l1p4s1,v0l0p3s1 > gpurun_out/r02d_ab.log 2> gpurun_out/r02d_ab.err; echo "ab rc=$?"
cut -c1-130 gpurun_out/r02d_ab.log
timeout 300 python bench.py --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r02d_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hybrid3 -s 12 -c 1 -o gpurun_out/r02d_h3 -f \
    python bench.py --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r02d_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r02d_ncu.log | cut -c1-200
