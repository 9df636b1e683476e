#!/bin/bash
# round 2, GPU call 3N (1 GPU): final tree -- whole parity suite, smoke, the driver's bench command (both arms)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r03n_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r03n_pytest.log
timeout 300 python __graft_entry__.py smoke
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03n_ref.json 2> gpurun_out/r03n_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03n_bench_k20.json 2> gpurun_out/r03n_bench_k20.err; echo "bench20 rc=$?"
python tools/brief.py k20 < gpurun_out/r03n_bench_k20.json
python - <<'PY'
import json
r=json.loads(open('gpurun_out/r03n_ref.json').read().strip().splitlines()[-1])
b=json.loads(open('gpurun_out/r03n_bench_k20.json').read().strip().splitlines()[-1])
print('reference %.4g  value ratio %.1f  e2e ratio %.1f' % (r['value'], b['value']/r['value'], b['e2e']['value']/r['value']))
print('riesz %.4f lbfgs %.4f legacy %.4f large %.4f sharded1 %.3f' % (b['riesz_gd']['ms_per_gd_step'], b['live_lbfgs']['ms_per_step'], b['live_lbfgs']['legacy_lbfgs']['ms_per_step'], b['large_n']['ms_per_bfgs_step'], b['sharded_large_n']['frac_of_peak']))
PY
