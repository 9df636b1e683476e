#!/bin/bash
# round 2, GPU call 3I (1 GPU): final tree -- whole parity suite, both bench lines, launch list, ncu of the Riesz kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r03i_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r03i_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r03i_bench_k20.json 2> gpurun_out/r03i_bench_k20.err; echo "bench20 rc=$?"
python tools/brief.py k20 < gpurun_out/r03i_bench_k20.json
timeout 600 python bench.py > gpurun_out/r03i_bench_default.json 2> gpurun_out/r03i_bench_default.err; echo "bench default rc=$?"
python tools/brief.py default < gpurun_out/r03i_bench_default.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r03i_launches.csv \
    python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/r03i_ncu_launches.log 2>&1; echo "launches rc=$?"
timeout 300 python tools/riesz_probe.py 4096 3 > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:riesz_gd_kernel -s 2 -c 1 -o gpurun_out/r03i_riesz -f python tools/riesz_probe.py 4096 3 > gpurun_out/r03i_ncu_riesz.log 2>&1; echo "ncu riesz rc=$?"
timeout 300 python __graft_entry__.py smoke
