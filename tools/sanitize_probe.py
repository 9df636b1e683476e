"""Small end-to-end run of every kernel family (target for compute-sanitizer)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dzopt_b200 as dz
import oracle as orc
EF = dz.ExampleFunctions
# batched hybrid (n=16, n=2) and first-generation kernel (n=6)
for n, batch in ((16, 70), (2, 100), (6, 40)):
    x0 = (4.0 * orc.pcg_fill(n * batch, 1) - 2.0).reshape(batch, n)
    o = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
    o.step(1); o.step(3)
    r = orc.BFGS(1, x0, 1.0); r.step(4)
    assert np.array_equal(o.current_point, r.point), n
    o.close()
# large path: cluster search + sweeps (n = 1100: two chunks, ragged tiles)
x0 = 4.0 * orc.pcg_fill(1100, 1) - 2.0
o = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0)
o.step(4)
r = orc.BFGS(1, x0[None, :], 1.0, order=orc.TREE); r.step(4)
assert np.array_equal(o.current_point, r.point[0])
o.close()
# GD Riesz cooperative kernel + Rosenbrock GD
p = 2.0 * orc.pcg_fill(3 * 150, 3).reshape(150, 3) - 1.0
p /= np.sqrt((p * p).sum(axis=1, keepdims=True))
g = dz.GradientDescentOptimizer(dz.SPHERE_CONSTRAINT, EF.riesz_energy, EF.riesz_gradient_, dz.QuadraticLineSearch(0), p, 1e-3)
g.step(2); g.close()
g = dz.GradientDescentOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, dz.QuadraticLineSearch(0), x0[:64], 1e-2)
g.step(3); g.close()
# pairwise
n = 300
c = orc.pcg_fill(3 * n, 5).reshape(3, n) * 8.0
for order in (0, 1):
    dz.accelerated_pairwise_radial_energy(EF.lj_energy, c[0], c[1], c[2], order=order)
    out = [np.empty(n) for _ in range(3)]
    dz.accelerated_pairwise_radial_gradient_(*out, EF.lj_first_derivative, c[0], c[1], c[2], order=order)
    dz.accelerated_pairwise_radial_hvp_(*out, EF.lj_first_derivative, EF.lj_second_derivative, c[0], c[1], c[2], c[0], c[1], c[2], order=order)
print("sanitize probe ok")
