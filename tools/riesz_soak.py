"""Soak test of the Riesz GD kernel's fused phases (completion counters, double-buffered rows, barrier-free hand-off):
long runs, k steps per launch, compared bit for bit with the oracle after every launch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dzopt_b200 as dz
import oracle as orc
EF = dz.ExampleFunctions
bad = 0
for N, launches, k, seed in ((4096, 12, 10, 3), (1000, 40, 13, 4), (333, 60, 7, 5), (129, 80, 5, 6)):
    p = 2.0 * dz.pcg_fill(3 * N, seed).reshape(N, 3) - 1.0
    p = p / np.sqrt((p * p).sum(axis=1, keepdims=True))
    o = dz.GradientDescentOptimizer(dz.SPHERE_CONSTRAINT, EF.riesz_energy, EF.riesz_gradient_, dz.QuadraticLineSearch(0), p, 1e-3)
    r = orc.GD(orc.OBJ_RIESZ, p.reshape(1, -1), 1e-3, order=orc.TREE, constraint=orc.CONSTRAINT_SPHERE, dim=3)
    for it in range(launches):
        o.step(k); r.step(k)
        same = (np.array_equal(o.current_point.reshape(-1), r.point[0]) and np.array_equal(o.next_step_direction.reshape(-1), r.direction[0])
                and float(o.current_objective_value[()]) == float(r.objective[0]) and int(o.iteration_count[()]) == int(r.iteration_count[0]))
        if not same:
            bad += 1
            print(f"MISMATCH N={N} launch {it}")
            break
    print(f"N={N}: {launches} launches x {k} steps, iterations {int(o.iteration_count[()])}, f={float(o.current_objective_value[()])!r}, ok={same}")
print("soak", "FAILED" if bad else "passed")
