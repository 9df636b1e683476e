#!/bin/bash
# round 2, GPU call I (2 GPUs): whole parity suite (incl. the row-sharded tests), bench.py under torchrun with 2 ranks, reference arm under torchrun
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02i_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02i_bench2.json 2> gpurun_out/r02i_bench2.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r02i_bench2.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/r02i_bench2.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','e2e','batched_strong','sharded_large_n','sharded_check'):
        print(k, json.dumps(b.get(k))[:900])
except Exception as e:
    print('parse failed', e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02i_ref2.json 2> gpurun_out/r02i_ref2.err; echo "ref2 rc=$?"
cut -c1-300 gpurun_out/r02i_ref2.json
timeout 600 python bench.py --steps 60 --warmup 5 --skip-large > gpurun_out/r02i_bench1_k60.json 2> gpurun_out/r02i_bench1_k60.err; echo "bench1 rc=$?"
python tools/brief.py k60 < gpurun_out/r02i_bench1_k60.json
