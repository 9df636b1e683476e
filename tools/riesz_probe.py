"""GD on Riesz N=4096 (config 5): time k step! calls; profiling target for riesz_gd_kernel."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dzopt_b200 as dz
import oracle as orc
EF = dz.ExampleFunctions
if os.environ.get('DZO_RIESZ_BAR'):
    dz.set_tuning('riesz_bar', int(os.environ['DZO_RIESZ_BAR']))
if os.environ.get('DZO_RIESZ_THREADS'):
    dz.set_tuning('riesz_threads', int(os.environ['DZO_RIESZ_THREADS']))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
if os.environ.get('DZO_RIESZ_ESPLIT'):
    dz.set_tuning('riesz_esplit', int(os.environ['DZO_RIESZ_ESPLIT']))
if os.environ.get('DZO_RIESZ_GVARIANT'):
    dz.set_tuning('riesz_gvariant', int(os.environ['DZO_RIESZ_GVARIANT']))
if os.environ.get('DZO_RIESZ_PAIR'):
    dz.set_tuning('riesz_pair', int(os.environ['DZO_RIESZ_PAIR']))
p = 2.0 * orc.pcg_fill(3 * N, 3).reshape(N, 3) - 1.0
p = p / np.sqrt((p * p).sum(axis=1, keepdims=True))
o = dz.GradientDescentOptimizer(dz.SPHERE_CONSTRAINT, EF.riesz_energy, EF.riesz_gradient_, dz.QuadraticLineSearch(0), p, 1e-3)
o.step(2)
for rep in range(3):
    it0 = int(o.iteration_count[()])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    o.step(k)
    dt = time.perf_counter() - t0
    print(f"{k} steps: {1e3*dt:.3f} ms -> {1e3*dt/k:.3f} ms/step, iterations {int(o.iteration_count[()])-it0}, f={float(o.current_objective_value[()])!r}")
