#!/bin/bash
# round 2, GPU call T (1 GPU): whole parity suite, default bench line, ncu captures of the Riesz and grid L-BFGS kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02t_pytest.log
( time timeout 900 python bench.py > gpurun_out/r02t_bench_default.json 2> gpurun_out/r02t_bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"
python tools/brief.py default < gpurun_out/r02t_bench_default.json
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02t_bench_k20.json 2> gpurun_out/r02t_bench_k20.err; echo "bench20 rc=$?"
python tools/brief.py k20 < gpurun_out/r02t_bench_k20.json
timeout 300 python tools/riesz_probe.py 4096 3 > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:riesz_gd_kernel -s 2 -c 1 -o gpurun_out/r02t_riesz -f python tools/riesz_probe.py 4096 3 > gpurun_out/r02t_ncu_riesz.log 2>&1; echo "ncu riesz rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_lbfgs_kernel -s 2 -c 1 -o gpurun_out/r02t_lbfgs -f python tools/lbfgs_probe.py 1048576 10 5 > gpurun_out/r02t_ncu_lbfgs.log 2>&1; echo "ncu lbfgs rc=$?"
