#!/bin/bash
# round 2, GPU call A (1 GPU): parity of the new batched kernel, A/B of the kernel generations, first bench line, ncu of the new kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python tools/batched_ab.py --configs v2l0p2,v0l0p2,v0l1p2,v0l1p3,v0l1p1,v0l1p4 > gpurun_out/r02a_ab.log 2> gpurun_out/r02a_ab.err; echo "ab rc=$?"
cat gpurun_out/r02a_ab.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02a_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r02a_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hybrid3 -s 12 -c 1 -o gpurun_out/r02a_h3 -f \
    python bench.py --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r02a_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r02a_ncu.log
