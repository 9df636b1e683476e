"""CPU tests (no GPU): the oracle against the golden fixtures, against the independent Python
restatement, against the run_and_test! invariants of the reference, SciPy and analytic minima."""
import json
import math
import os

import numpy as np
import pytest

from conftest import assert_bitwise

HERE = os.path.dirname(os.path.abspath(__file__))
ROSEN, RIESZ = 1, 2
fh = float.fromhex


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        return json.load(f)


def unhex(v):
    return np.array([fh(a) for a in v])


# ----------------------------------------------------------------------------- golden vectors
def test_pcg_golden(orc, golden):
    for seed, vals in golden["pcg"].items():
        assert_bitwise(orc.pcg_fill(8, int(seed)), unhex(vals), f"pcg seed {seed}")
    u = orc.pcg_fill(100000, 5)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01   # legacy/PCG.jl:18 uniform [0,1)


def _replay_bfgs(orc, case, order):
    x0 = unhex(case["x0"])
    opt = orc.BFGS(ROSEN, x0[None, :], case["step"], order=order)
    for k, row in enumerate(case["rows"]):
        opt.step(1)
        assert opt.objective[0] == fh(row["f"]), f"iter {k}: objective"
        assert opt.step_length[0] == fh(row["L"]), f"iter {k}: step length"
        assert int(opt.step_type[0]) == row["type"] and int(opt.iteration_count[0]) == row["iter"]
        assert bool(opt.terminated[0]) == row["term"]
        if "x" in row:
            assert_bitwise(opt.point[0], unhex(row["x"]), f"iter {k}: point")
    assert_bitwise(opt.point[0], unhex(case["final_x"]), "final point")
    assert_bitwise(opt.direction[0], unhex(case["final_d"]), "final direction")
    assert_bitwise(opt.inverse_hessian(0)[0], unhex(case["final_H_row0"]), "H row 0")
    return opt


def test_bfgs_golden_c1_readme_rosenbrock(orc, golden):
    for case in golden["c1_rosenbrock_n2"]:
        opt = _replay_bfgs(orc, case, orc.SEQ)
        assert case["rows"][-1]["term"]                       # ran to has_converged
        assert np.abs(opt.point[0] - 1.0).max() < 1e-6        # README.md:33-41 converges to (1, 1)


def test_bfgs_golden_c2_n16(orc, golden):
    for case in golden["c2_rosenbrock_n16"]:
        _replay_bfgs(orc, case, orc.SEQ)


def test_bfgs_golden_tree(orc, golden):
    _replay_bfgs(orc, golden["tree_rosenbrock_n64"], orc.TREE)
    _replay_bfgs(orc, golden["tree_rosenbrock_n1100"], orc.TREE)


def test_tree_kernels_golden(orc, golden):
    n = 1030
    a = orc.pcg_fill(3 * n, 77)
    assert orc.dot(a[:n], a[n:2 * n], orc.TREE) == fh(golden["tree_dot_n1030"]["value"])
    rows = golden["tree_gemv_rows_n1030"]["rows"]
    H = np.zeros((n, n))
    H[:rows] = orc.pcg_fill(n * rows, 78).reshape(rows, n) - 0.5
    out = orc.gemv(H, a[2 * n:], orc.TREE)
    assert_bitwise(out[:rows], unhex(golden["tree_gemv_rows_n1030"]["values"]), "tree gemv rows")


def _sphere_points(orc, N, dim, seed):
    p = 2.0 * orc.pcg_fill(N * dim, seed).reshape(N, dim) - 1.0
    return p / np.sqrt((p * p).sum(axis=1, keepdims=True))


def _replay_gd(orc, case, obj, order, constraint=0, dim=0):
    x0 = unhex(case["x0"])
    opt = orc.GD(obj, x0[None, :], case["step"], order=order, constraint=constraint, dim=dim,
                 max_increases=case["max_increases"])
    for k, row in enumerate(case["rows"]):
        opt.step(1)
        assert opt.objective[0] == fh(row["f"]), f"iter {k}: objective"
        assert opt.step_length[0] == fh(row["L"]), f"iter {k}: step length"
        assert int(opt.iteration_count[0]) == row["iter"] and bool(opt.terminated[0]) == row["term"]
    assert_bitwise(opt.point[0], unhex(case["final_x"]), "final point")
    assert_bitwise(opt.direction[0], unhex(case["final_d"]), "final direction")


def test_gd_golden(orc, golden):
    _replay_gd(orc, golden["gd_riesz_sphere_N40"], RIESZ, orc.TREE, 1, 3)
    _replay_gd(orc, golden["gd_riesz_sphere_N300"], RIESZ, orc.TREE, 1, 3)
    _replay_gd(orc, golden["gd_riesz_free_N20_seq"], RIESZ, orc.SEQ, 0, 2)
    _replay_gd(orc, golden["gd_rosenbrock_n64"], ROSEN, orc.TREE)


def test_riesz_golden(orc, golden):
    g = golden["riesz_N300"]
    # the fixture's points come from the Python generator; rebuild them the same way
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    pts = np.array(make_golden.sphere_points(300, 3, g["seed"]))
    assert orc.objective(RIESZ, pts, orc.TREE, 1, 3)[0] == fh(g["energy_tree"])
    assert orc.objective(RIESZ, pts, orc.SEQ, 1, 3)[0] == fh(g["energy_seq"])
    assert_bitwise(orc.gradient(RIESZ, pts, orc.TREE, 1, 3)[0][:6], unhex(g["gradient_tree_first6"]), "gradient")


# ----------------------------------------------------------------------------- C oracle vs Python restatement
@pytest.mark.parametrize("n,tree,seed", [(2, False, 1), (6, False, 2), (10, True, 3), (40, True, 4)])
def test_c_oracle_equals_python_restatement_bfgs(orc, n, tree, seed):
    import dzo_oracle_py as P
    x0 = (4.0 * orc.pcg_fill(n, seed) - 2.0)
    py = P.BFGSOptimizer(P.Rosenbrock(tree), list(x0), 0.5, tree)
    c = orc.BFGS(ROSEN, x0[None, :], 0.5, order=orc.TREE if tree else orc.SEQ)
    for it in range(30):
        py.step(); c.step(1)
        assert_bitwise(c.point[0], np.array(py.current_point), f"iter {it} point")
        assert_bitwise(c.gradient[0], np.array(py.current_gradient), f"iter {it} gradient")
        assert_bitwise(c.direction[0], np.array(py.next_step_direction), f"iter {it} direction")
        assert_bitwise(c.inverse_hessian(0), np.array(py.H), f"iter {it} H")
        assert c.objective[0] == py.current_objective_value and c.step_length[0] == py.last_step_length
        assert bool(c.terminated[0]) == py.has_terminated


@pytest.mark.parametrize("N,dim,sphere,tree", [(7, 3, True, False), (150, 3, True, True), (20, 2, False, True)])
def test_c_oracle_equals_python_restatement_gd_riesz(orc, N, dim, sphere, tree):
    import dzo_oracle_py as P
    x0 = _sphere_points(orc, N, dim, 9).reshape(-1)
    py = P.GradientDescentOptimizer(P.Riesz(dim, sphere, tree), list(x0), 1e-2, 0, tree)
    c = orc.GD(RIESZ, x0[None, :], 1e-2, order=orc.TREE if tree else orc.SEQ, constraint=int(sphere), dim=dim)
    for it in range(6):
        py.step(); c.step(1)
        assert_bitwise(c.point[0], np.array(py.current_point), f"iter {it} point")
        assert_bitwise(c.direction[0], np.array(py.next_step_direction), f"iter {it} direction")
        assert c.objective[0] == py.current_objective_value and c.step_length[0] == py.last_step_length


def test_line_search_matches_python_restatement(orc):
    import dzo_oracle_py as P
    for n, t1, tree in ((2, 1.0, False), (16, 1e-4, False), (16, 1e4, False), (64, 3.0, True), (64, 1e-200, True)):
        x = 4.0 * orc.pcg_fill(n, 31) - 2.0
        fn = P.Rosenbrock(tree)
        g = [0.0] * n
        fn.g(g, list(x))
        f0 = fn.f(list(x))
        ref = P.quadratic_line_search(P.Ray(fn, list(x), g, -1.0), f0, t1, 0)
        assert orc.line_search(ROSEN, x, np.array(g), f0, t1, orc.TREE if tree else orc.SEQ) == ref


# ----------------------------------------------------------------------------- run_and_test! invariants
def test_run_and_test_invariants(orc):
    """legacy/DZOptimization.jl:998-1049, on the oracle."""
    n = 8
    x0 = 4.0 * orc.pcg_fill(n, 12) - 2.0
    opt = orc.BFGS(ROSEN, x0[None, :], 1.0)
    hist = [(opt.point[0], opt.gradient[0], 0, False)]
    for _ in range(2000):
        opt.step(1)
        hist.append((opt.point[0], opt.gradient[0], int(opt.iteration_count[0]), bool(opt.terminated[0])))
        assert opt.objective[0] == orc.objective(ROSEN, opt.point[0])[0]                       # :1019-1022
        assert_bitwise(orc.gradient(ROSEN, opt.point[0])[0], opt.gradient[0], "gradient")      # :1025-1032
        if not hist[-1][3]:
            assert_bitwise(opt.delta_point[0], (-hist[-2][0]) + hist[-1][0], "delta_point")    # :1035-1039
            assert_bitwise(opt.delta_gradient[0], (-hist[-2][1]) + hist[-1][1], "delta_grad")  # :1042-1046
        if hist[-1][3]:
            break
    assert hist[-1][3] and not any(h[3] for h in hist[:-1])                                    # :1007-1010
    assert [h[2] for h in hist[:-1]] == list(range(len(hist) - 1))                             # :1013-1016
    assert_bitwise(hist[-1][0], hist[-2][0], "final step leaves the point unchanged")          # :1039
    assert np.abs(hist[-1][0] - 1.0).max() < 1e-6


def test_finite_difference_gradients(orc):
    """finite_difference_gradient, legacy/ExampleFunctions.jl:290-303 (central differences)."""
    def fd(f, x, h):
        g = np.zeros_like(x)
        for i in range(x.size):
            xp, xm = x.copy(), x.copy()
            xp[i] += h; xm[i] -= h
            g[i] = (f(xp) - f(xm)) / (2 * h)
        return g
    x = 4.0 * orc.pcg_fill(10, 2) - 2.0
    for order in (orc.SEQ, orc.TREE):
        g = orc.gradient(ROSEN, x, order)[0]
        assert np.allclose(g, fd(lambda v: orc.objective(ROSEN, v, order)[0], x, 1e-6), rtol=1e-6, atol=1e-5)
    p = _sphere_points(orc, 12, 3, 3).reshape(-1) * 1.3
    for order in (orc.SEQ, orc.TREE):
        g = orc.gradient(RIESZ, p, order, 0, 3)[0]
        assert np.allclose(g, fd(lambda v: orc.objective(RIESZ, v, order, 0, 3)[0], p, 1e-6), rtol=1e-5, atol=1e-5)
    # the projected gradient is tangent to the sphere
    q = _sphere_points(orc, 12, 3, 3)
    gt = orc.gradient(RIESZ, q.reshape(-1), orc.SEQ, 1, 3)[0].reshape(12, 3)
    assert np.abs((gt * q).sum(axis=1)).max() < 1e-12


def test_scipy_cross_check(orc):
    """Independent check of the converged point (iterates differ: SciPy uses a Wolfe line search)."""
    from scipy.optimize import minimize, rosen, rosen_der  # noqa: F401  (classic chained form differs from ours)
    n = 6
    x0 = 2.0 * orc.pcg_fill(n, 4)
    f = lambda v: orc.objective(ROSEN, v)[0]
    g = lambda v: orc.gradient(ROSEN, v)[0]
    res = minimize(f, x0, jac=g, method="BFGS", options={"gtol": 1e-10})
    opt = orc.BFGS(ROSEN, x0[None, :], 1.0)
    for _ in range(100):
        opt.step(50)
        if opt.terminated[0]:
            break
    assert opt.terminated[0]
    assert np.abs(opt.point[0] - res.x).max() < 1e-6
    assert np.abs(opt.point[0] - 1.0).max() < 1e-6 and opt.objective[0] < 1e-12


def test_tree_and_sequential_orders_agree_to_rounding(orc):
    """north_star tolerance: the two summation orders of one algorithm agree to ~1e-12 relative for the
    first iterations (until a branch flips; the harness reports where)."""
    n = 64
    x0 = 4.0 * orc.pcg_fill(n, 1) - 2.0
    a = orc.BFGS(ROSEN, x0[None, :], 1.0, order=orc.SEQ)
    b = orc.BFGS(ROSEN, x0[None, :], 1.0, order=orc.TREE)
    agree = 0
    for _ in range(10):
        a.step(1); b.step(1)
        if a.step_type[0] != b.step_type[0] or abs(a.step_length[0] - b.step_length[0]) > 1e-9 * abs(a.step_length[0]):
            break
        agree += 1
        assert abs(a.objective[0] - b.objective[0]) <= 1e-12 * abs(a.objective[0])
        assert np.abs(a.point[0] - b.point[0]).max() <= 1e-12 * np.abs(a.point[0]).max()
    assert agree >= 3


def test_errors_and_edge_cases(orc):
    with pytest.raises(orc.OracleError) as e:
        orc.BFGS(ROSEN, np.array([[np.nan, 1.0]]), 1.0)
    assert e.value.code == -3                                        # @assert !isnan  :773
    with pytest.raises(orc.OracleError):
        orc.BFGS(ROSEN, np.zeros((1, 3)), 1.0)                       # odd n
    # already at the minimum: gradient 0 -> both searches return step 0 -> terminated, nothing moves
    o = orc.BFGS(ROSEN, np.ones((1, 4)), 1.0)
    o.step(1)
    assert o.terminated[0] and o.iteration_count[0] == 0 and (o.point[0] == 1.0).all()
    # GD from a non-finite start: state, not an error  :364-366
    g = orc.GD(RIESZ, np.array([[0.0, 0.0, 1.0, 0.0, 0.0, 1.0]]), 1e-2, constraint=1, dim=3)
    assert g.terminated[0]
    # zero initial step length: L/||g|| = 0 -> no probe -> terminates immediately [GLUE]
    z = orc.BFGS(ROSEN, np.zeros((1, 2)), 0.0)
    z.step(1)
    assert z.terminated[0]
    assert math.isfinite(z.objective[0])


@pytest.mark.parametrize("n,order_name", [(16, "SEQ"), (2048, "TREE")])
def test_secant_condition_after_every_bfgs_step(orc, n, order_name):
    """A check of update_inverse_hessian! (legacy/DZOptimization.jl:864-889) that neither restatement produced: the
    BFGS update is DEFINED by the secant condition H_new * delta_gradient = delta_point.  After every BFGS-type step
    the oracle's inverse Hessian satisfies it to rounding (observed 2e-14 ... 5e-13 relative) and is bitwise symmetric.
    The CUDA twin of this test is tests/test_gpu_bfgs.py::test_secant_condition_after_every_bfgs_step."""
    order = getattr(orc, order_name)
    batch = 64 if n == 16 else 1
    x0 = (4.0 * orc.pcg_fill(n * batch, 612) - 2.0).reshape(batch, n)
    opt = orc.BFGS(ROSEN, x0, 1.0, order=order, nthreads=4)
    seen, worst = 0, 0.0
    for it in range(25 if n == 16 else 12):
        opt.step(1)
        ty, moved = opt.step_type, opt.iteration_count == it + 1
        dx, dg = opt.delta_point, opt.delta_gradient
        for p in range(batch):
            if not (moved[p] and ty[p] == 2):
                continue
            H = opt.inverse_hessian(p)
            assert_bitwise(H, H.T, f"iter {it} problem {p}: H symmetric")
            res = float(np.abs(H @ dg[p] - dx[p]).max() / np.abs(dx[p]).max())
            worst = max(worst, res)
            assert res < 1e-10, f"iter {it} problem {p}: secant condition violated ({res:.3e})"
            seen += 1
    assert seen > (100 if n == 16 else 5)
    assert worst > 0.0      # it is a floating-point identity, not an algebraic coincidence of the test


def test_native_build_of_the_oracle_is_bit_identical(orc):
    """bench.py times the oracle compiled with -O3 -march=native (a fairer CPU arm); no contraction, no reassociation:
    the trace must be the -O2 build's, bit for bit."""
    x0 = (4.0 * orc.pcg_fill(16 * 300, 2024) - 2.0).reshape(300, 16)
    a = orc.BFGS(ROSEN, x0, 1.0, order=orc.SEQ, nthreads=2)
    a.step(30)
    pa, fa, Ha = a.point, a.objective, a.inverse_hessian(7)
    xl = (4.0 * orc.pcg_fill(1500, 1) - 2.0)
    al = orc.BFGS(ROSEN, xl[None, :], 1.0, order=orc.TREE, nthreads=2)
    al.step(6)
    pl, dl = al.point, al.direction
    a.close(); al.close()
    try:
        flags = orc.use_native_build()
        assert "-march=native" in flags and "-ffp-contract=off" in flags
        b = orc.BFGS(ROSEN, x0, 1.0, order=orc.SEQ, nthreads=2)
        b.step(30)
        assert_bitwise(b.point, pa, "native build: batched point")
        assert_bitwise(b.objective, fa, "native build: objective")
        assert_bitwise(b.inverse_hessian(7), Ha, "native build: H")
        bl = orc.BFGS(ROSEN, xl[None, :], 1.0, order=orc.TREE, nthreads=2)
        bl.step(6)
        assert_bitwise(bl.point, pl, "native build: large-n point")
        assert_bitwise(bl.direction, dl, "native build: large-n direction")
        b.close(); bl.close()
    finally:
        orc.use_default_build()
