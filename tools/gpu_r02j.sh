#!/bin/bash
# round 2, GPU call J (1 GPU): paired probe evaluation of the Riesz line search -- parity, A/B, phase log; bench.py over three batches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gd.py -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02j_pytest.log
for pair in 0 1 0 1; do echo "== riesz_pair=$pair"; DZO_RIESZ_PAIR=$pair timeout 300 python tools/riesz_probe.py 4096 20 | tail -2; done 2>&1 | tee gpurun_out/r02j_riesz_ab.log
for pair in 0 1; do echo "== phases riesz_pair=$pair"; DZO_RIESZ_PAIR=$pair timeout 300 python tools/riesz_phases.py 4096 20; done > gpurun_out/r02j_riesz_phases.log 2>&1
cat gpurun_out/r02j_riesz_phases.log
timeout 600 python bench.py --steps 60 --warmup 5 --skip-large --skip-cpu > gpurun_out/r02j_bench_k60.json 2> gpurun_out/r02j_bench_k60.err; echo "bench rc=$?"
python tools/brief.py k60 < gpurun_out/r02j_bench_k60.json
python -c "
import json
b=json.loads(open('gpurun_out/r02j_bench_k60.json').read().strip().splitlines()[-1])
print(json.dumps(b['e2e'])[:1200])"
