"""legacy LBFGSOptimizer at large n: ms per step!."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dzopt_b200 as dz
import oracle as orc
EF = dz.ExampleFunctions
if os.environ.get('DZO_GRID_STAGE'):
    dz.set_tuning('grid_stage', int(os.environ['DZO_GRID_STAGE']))
if os.environ.get('DZO_GRID_BACKOFF'):
    dz.set_tuning('grid_ll_backoff', int(os.environ['DZO_GRID_BACKOFF']))
if os.environ.get('DZO_GRID_LL'):
    dz.set_tuning('grid_ll', int(os.environ['DZO_GRID_LL']))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10
k = int(sys.argv[3]) if len(sys.argv) > 3 else 50
x0 = 4.0 * orc.pcg_fill(n, 1) - 2.0
o = dz.LegacyLBFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, dz.QuadraticLineSearch(0), x0, 1.0, m)
o.step(12)
for rep in range(3):
    it0 = int(o.iteration_count[()])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    o.step(k)
    dt = time.perf_counter() - t0
    print(f"n={n} m={m}: {k} steps in one launch: {1e3*dt/k:.4f} ms/step  f={float(o.current_objective_value[()])!r} iterations {int(o.iteration_count[()])-it0}")
