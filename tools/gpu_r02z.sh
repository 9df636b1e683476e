#!/bin/bash
# round 2, GPU call Z (1 GPU): final grid-wide reduction code -- parity of every user, bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lbfgs.py tests/test_legacy_lbfgs.py tests/test_gpu_gd.py -m gpu -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench_k20.json 2> gpurun_out/r02z_bench_k20.err; echo "bench20 rc=$?"
python tools/brief.py k20 < gpurun_out/r02z_bench_k20.json
python -c "
import json
b=json.loads(open('gpurun_out/r02z_bench_k20.json').read().strip().splitlines()[-1])
print(json.dumps(b['live_lbfgs'])[:900])"
DZO_GRID_PROFILE=1 timeout 120 python tools/lbfgs_probe.py 1048576 10 50 | tail -3
