#!/bin/bash
# round 2, GPU call F (1 GPU): consolidated batched kernel for every even n <= 32 -- full parity suite, per-n roofline lines, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02f_pytest.log
for n in 16 6 12 24 32; do
  timeout 600 python tools/batched_ab.py --n $n --batch $((16000000 / n)) --configs l1p3,l0p3 --repeat 1 >> gpurun_out/r02f_ab.log 2>> gpurun_out/r02f_ab.err; echo "ab n=$n rc=$?"
done
cut -c1-420 gpurun_out/r02f_ab.log
timeout 300 python bench.py --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r02f_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:batched_hybrid -s 12 -c 1 -o gpurun_out/r02f_hybrid -f \
    python bench.py --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r02f_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r02f_ncu.log | cut -c1-200
