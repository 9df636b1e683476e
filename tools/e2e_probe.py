"""Where does the end-to-end time of the batched README loop go? (constructor / step! / field reads)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dzopt_b200 as dz
import oracle as orc
EF = dz.ExampleFunctions
x0 = (4.0 * orc.pcg_fill(16 * 1_000_000, 2024) - 2.0).reshape(-1, 16)
xp = torch.from_numpy(x0).pin_memory().numpy()
for rep in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, xp, 1.0, batched=True)
    t1 = time.perf_counter()
    o.reuse_host_buffers(True)
    ts, tr, worst = 0.0, 0.0, (0.0, 0.0)
    for _ in range(25):
        a = time.perf_counter(); o.step(1); b = time.perf_counter()
        f = o.has_converged; v = o.current_objective_value; c = time.perf_counter()
        ts += b - a; tr += c - b
        worst = (max(worst[0], b - a), max(worst[1], c - b))
    t2 = time.perf_counter()
    o.close()
    t3 = time.perf_counter()
    print(f"rep {rep}: ctor {1e3*(t1-t0):.1f} ms  25 steps {1e3*ts:.1f} ms  25 reads {1e3*tr:.1f} ms  close {1e3*(t3-t2):.1f} ms  worst step {1e3*worst[0]:.2f} ms  worst read {1e3*worst[1]:.2f} ms")
