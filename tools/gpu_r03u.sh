#!/bin/bash
# round 2, GPU call 3U (1 GPU): warp-per-problem n^2 sweeps for batches of n <= 128 -- parity, A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bfgs.py -m gpu -x -q -k "medium or warp" > gpurun_out/r03u_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r03u_pytest.log
python - <<'PY' 2>&1 | tee gpurun_out/r03u_small_sweeps.jsonl
import json, sys
sys.path.insert(0, '.')
import torch, bench
import dzopt_b200 as dz
EF = dz.ExampleFunctions
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for n, batch in ((34, 16384), (64, 16384), (100, 8192), (128, 8192)):
    x0 = (4.0 * dz.pcg_fill(n * batch, 7) - 2.0).reshape(batch, n)
    for ss in (0, 1, 0, 1):
        dz.set_tuning("small_sweeps", ss)
        opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
        opt.set_stream(stream.cuda_stream)
        opt.step(3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): opt.step_async(1)
        e1.record(stream); torch.cuda.synchronize()
        print(json.dumps({"n": n, "batch": batch, "small_sweeps": ss, "ms_per_step_call": e0.elapsed_time(e1) / 10}), flush=True)
        opt.close()
dz.set_tuning("small_sweeps", 1)
PY
