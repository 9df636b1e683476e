"""n = 16384 single-problem BFGS: a few step! calls (profiling target for the large-n kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dzopt_b200 as dz
import oracle as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
EF = dz.ExampleFunctions
x0 = 4.0 * orc.pcg_fill(n, 1) - 2.0
o = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0)
s = torch.cuda.Stream(); o.set_stream(s.cuda_stream)
o.step(3)
ts = []
for _ in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); o.step_async(1); e1.record(s); torch.cuda.synchronize()
    ts.append((round(e0.elapsed_time(e1), 4), int(o.last_step_type[()])))
print(ts)
