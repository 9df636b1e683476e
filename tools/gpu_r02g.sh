#!/bin/bash
# round 2, GPU call G (1 GPU): batch of medium-n problems (one cluster per problem); kind-aware L2 prefetch of the batched kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bfgs.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02g_pytest.log
timeout 300 python tools/medium_probe.py > gpurun_out/r02g_medium.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r02g_medium.log | tail -12
timeout 600 python tools/batched_ab.py --configs l1p3,l1p2,l1p4,l0p3 > gpurun_out/r02g_ab.log 2> gpurun_out/r02g_ab.err; echo "ab rc=$?"
cut -c1-130 gpurun_out/r02g_ab.log
timeout 300 python bench.py --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r02g_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:batched_hybrid -s 12 -c 1 -o gpurun_out/r02g_hybrid -f \
    python bench.py --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r02g_ncu.log 2>&1; echo "ncu rc=$?"
