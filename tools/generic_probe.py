"""Throughput of the one-thread-per-problem batched kernels (GradientDescentOptimizer n <= 32, BFGSOptimizer with the Riesz objective)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dzopt_b200 as dz
EF = dz.ExampleFunctions


def timed(opt, steps=10):
    opt.step(3)
    it0 = opt.iteration_count.copy()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        opt.step(1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    moved = float((opt.iteration_count - it0).sum()) / steps
    return ms, moved


batch = 200_000
x0 = (4.0 * dz.pcg_fill(batch * 16, 5) - 2.0).reshape(batch, 16)
gd = dz.GradientDescentOptimizer(dz.NULL_CONSTRAINT, EF.rosenbrock_function, EF.rosenbrock_gradient_, dz.QuadraticLineSearch(0), x0, 1e-3, batched=True)
ms, moved = timed(gd)
print(json.dumps({"kernel": "gd_batched (Rosenbrock n=16)", "batch": batch, "ms_per_step": ms, "problem_steps_per_s": moved / (ms * 1e-3)}))
p = 2.0 * dz.pcg_fill(batch * 24, 6).reshape(batch, 8, 3) - 1.0
p /= np.sqrt((p * p).sum(axis=2, keepdims=True))
gr = dz.GradientDescentOptimizer(dz.SPHERE_CONSTRAINT, EF.riesz_energy, EF.riesz_gradient_, dz.QuadraticLineSearch(0), p, 1e-3, batched=True)
ms, moved = timed(gr)
print(json.dumps({"kernel": "gd_batched (Riesz 8 points x 3)", "batch": batch, "ms_per_step": ms, "problem_steps_per_s": moved / (ms * 1e-3)}))
bf = dz.BFGSOptimizer(EF.riesz_energy, EF.riesz_gradient_, dz.SPHERE_CONSTRAINT, p, 1e-3, batched=True)
ms, moved = timed(bf)
print(json.dumps({"kernel": "bfgs_generic (Riesz 8 points x 3)", "batch": batch, "ms_per_step": ms, "problem_steps_per_s": moved / (ms * 1e-3)}))
