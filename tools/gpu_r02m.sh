#!/bin/bash
# round 2, GPU call M (8 GPUs): row-sharded tests, bench.py --gpus 8 under torchrun (both arms)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02m_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02m_bench8.json 2> gpurun_out/r02m_bench8.err; echo "bench8 rc=$?"
tail -3 gpurun_out/r02m_bench8.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/r02m_bench8.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','e2e','batched_strong','sharded_large_n','sharded_check'):
        print(k, json.dumps(b.get(k))[:1500])
except Exception as e:
    print('parse failed', e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02m_ref8.json 2> gpurun_out/r02m_ref8.err; echo "ref8 rc=$?"
cut -c1-300 gpurun_out/r02m_ref8.json
