#!/bin/bash
# round 2, GPU call 3T (1 GPU): roofline line of the batched kernel for every tile shape (n = 2 ... 32, 16 M doubles of vectors per batch)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for n in 2 4 6 8 12 16 24 32; do
  timeout 300 python tools/batched_ab.py --n $n --batch $((16000000 / n)) --configs l1p3 --repeat 1 2>/dev/null
done | tee gpurun_out/r03t_batched_per_n.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); k = d['kinds']
    print(k.get('bfgs_read_h',0)+k.get('bfgs_identity_h',0)+k.get('gradient_descent',0), round(d['ms_per_step'],4), round(d['achieved_gbs']), round(d['frac'],3))"
