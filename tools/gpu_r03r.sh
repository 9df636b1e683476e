#!/bin/bash
# round 2, GPU call 3R (1 GPU): threads per sweep CTA for batches of n = 64 / 128 problems
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/r03r_medium_sweep_threads.jsonl
import json, sys
sys.path.insert(0, '.')
import torch, bench
import dzopt_b200 as dz
EF = dz.ExampleFunctions
peak, _ = bench.load_peaks()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for n, batch in ((64, 16384), (128, 8192), (256, 2048)):
    x0 = (4.0 * dz.pcg_fill(n * batch, 7) - 2.0).reshape(batch, n)
    for st in (0, 32, 64, 128):
        dz.set_tuning("sweep_threads", st)
        opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
        opt.set_stream(stream.cuda_stream)
        opt.step(3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): opt.step_async(1)
        e1.record(stream); torch.cuda.synchronize()
        print(json.dumps({"n": n, "batch": batch, "sweep_threads": st, "ms_per_step_call": e0.elapsed_time(e1) / 10}), flush=True)
        opt.close()
dz.set_tuning("sweep_threads", 0)
PY
