#!/bin/bash
# round 2, GPU call 3W (8 GPUs): bench.py --gpus 8 of the final tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r03w_bench8.json 2> gpurun_out/r03w_bench8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r03w_bench8.json').read().strip().splitlines()[-1])
print('value %.4g e2e %.4g frac %.3f' % (b['value'], b['e2e']['value'], b['roofline']['frac']))
s=b['batched_strong']; print('strong ms %.4f frac %.3f total %.4g e2e %.4g' % (s['ms_per_step'], s['per_gpu_frac_of_peak'], s['problem_steps_per_s'], s['e2e_problem_steps_per_s']))
h=b['sharded_large_n']
for m in ('fused-peer-memory','nccl-allgather'): print(m, round(h[m]['ms_per_bfgs_step'],4), round(h[m]['frac_of_peak'],4), h[m]['bitwise_check'])
PY
