#!/bin/bash
# round 2, GPU call B (1 GPU): parity of the reworked batched kernel, A/B, full bench line, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02b_pytest.log
timeout 600 python tools/batched_ab.py --configs v2l0p2,v0l0p2,v0l1p2,v0l1p3,v0l1p4,v0l1p2d1 > gpurun_out/r02b_ab.log 2> gpurun_out/r02b_ab.err; echo "ab rc=$?"
cut -c1-330 gpurun_out/r02b_ab.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02b_bench.err
cut -c1-600 gpurun_out/r02b_bench.json
timeout 300 python bench.py --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r02b_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hybrid3 -s 12 -c 1 -o gpurun_out/r02b_h3 -f \
    python bench.py --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r02b_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r02b_ncu.log | cut -c1-200
