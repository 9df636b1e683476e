#!/bin/bash
# round 2, GPU call 3B (2 GPUs): final tree -- whole parity suite incl. the row-sharded tests, bench.py under torchrun with 2 ranks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r03b_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r03b_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r03b_bench2.json 2> gpurun_out/r03b_bench2.err; echo "bench2 rc=$?"
python tools/brief.py n2 < gpurun_out/r03b_bench2.json
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r03b_bench2.json').read().strip().splitlines()[-1])
for k in ('batched_strong','sharded_large_n','sharded_check'):
    print(k, json.dumps(b.get(k))[:700])
print('e2e', b['e2e']['value'], b['e2e']['host_phases_ms'])
PY
