#!/bin/bash
# round 2, GPU call 3G (2 GPUs): warp-uniform peer_wait -- row-sharded tests, sharded bench leg
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r03g_bench2.json 2> gpurun_out/r03g_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r03g_bench2.json').read().strip().splitlines()[-1])
s=b['sharded_large_n']
print('value %.4g e2e %.4g' % (b['value'], b['e2e']['value']))
for m in ('fused-peer-memory','nccl-allgather'):
    print(m, s[m]['ms_per_bfgs_step'], s[m]['frac_of_peak'], s[m]['bitwise_check'])
PY
