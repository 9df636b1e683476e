"""Live LBFGSOptimizer (src/DZOptimization.jl:321-509) + take_backtracking_step! (:107-154):
oracle vs the independent Python restatement (CPU) and CUDA vs oracle (GPU), all bitwise."""
import numpy as np
import pytest

from conftest import assert_bitwise

ROSEN = 1


def _x0(orc, n, seed):
    return 4.0 * orc.pcg_fill(n, seed) - 2.0


@pytest.mark.parametrize("n,m,tree", [(2, 3, False), (10, 4, True), (64, 5, True), (30, 1, True)])
def test_c_oracle_equals_python_restatement(orc, n, m, tree):
    import dzo_oracle_py as P
    x0 = _x0(orc, n, 3)
    py = P.LiveLBFGSOptimizer(P.Rosenbrock(tree), list(x0), 0.5, m, tree)
    c = orc.LBFGS(ROSEN, x0, 0.5, m, orc.TREE if tree else orc.SEQ)
    for it in range(40):
        py.step(); c.step(1)
        assert_bitwise(c.point, np.array(py.current_point), f"iter {it} point")
        assert_bitwise(c.direction, np.array(py.step_direction), f"iter {it} direction")
        assert_bitwise(c.delta_point, np.array(py.delta_point), f"iter {it} delta_point")
        assert_bitwise(c.delta_gradient, np.array(py.delta_gradient), f"iter {it} delta_gradient")
        assert c.objective == py.current_objective_value and c.delta_objective == py.delta_objective_value
        assert c.iteration_count == py.iteration_count and c.stuck == py.is_stuck
        assert_bitwise(c.rho_history, np.array(py.rho), "rho history")


def test_lbfgs_converges_to_the_rosenbrock_minimum(orc):
    x0 = _x0(orc, 8, 4)
    c = orc.LBFGS(ROSEN, x0, 1.0, 6, orc.SEQ)
    for _ in range(400):
        c.step(50)
        if c.stuck:
            break
    assert c.stuck                                     # the live spelling of has_terminated
    assert np.abs(c.point - 1.0).max() < 1e-6 and c.objective < 1e-12
    # at the minimum the gradient is zero: constructor reports is_stuck at once  (:377)
    z = orc.LBFGS(ROSEN, np.ones(4), 1.0, 3)
    assert z.stuck and z.iteration_count == 0


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(2, 2), (34, 5), (2048, 8), (16384, 10), (20000, 3)])
def test_gpu_lbfgs_trace(gpu, orc, n, m):
    dz = gpu
    EF = dz.ExampleFunctions
    x0 = _x0(orc, n, 5)
    opt = dz.LBFGSOptimizer(None, EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, m)
    ref = orc.LBFGS(ROSEN, x0, 1.0, m, orc.TREE)

    def compare(tag):
        assert_bitwise(opt.current_point, ref.point, f"{tag}: point")
        assert_bitwise(opt.delta_point, ref.delta_point, f"{tag}: delta_point")
        assert_bitwise(opt.current_gradient, ref.gradient, f"{tag}: gradient")
        assert_bitwise(opt.delta_gradient, ref.delta_gradient, f"{tag}: delta_gradient")
        assert_bitwise(opt.step_direction, ref.direction, f"{tag}: direction")
        assert float(opt.current_objective_value[()]) == ref.objective
        assert float(opt.delta_objective_value[()]) == ref.delta_objective
        assert int(opt.iteration_count[()]) == ref.iteration_count and bool(opt.is_stuck[()]) == ref.stuck
        assert_bitwise(opt.rho_history, ref.rho_history, f"{tag}: rho")

    compare("ctor")
    for it in range(20):
        dz.step_(opt); ref.step(1)
        compare(f"n={n} iter {it}")
    opt.step(25); ref.step(25)            # 25 step! calls in one cluster-kernel launch
    compare("fused")


@pytest.mark.gpu
def test_gpu_lbfgs_to_convergence(gpu, orc):
    dz = gpu
    EF = dz.ExampleFunctions
    x0 = _x0(orc, 64, 6)
    opt = dz.LBFGSOptimizer(dz.NULL_CONSTRAINT, EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, 8)
    ref = orc.LBFGS(ROSEN, x0, 1.0, 8, orc.TREE)
    for _ in range(200):
        opt.step(100); ref.step(100)
        if opt.is_stuck[()]:
            break
    assert bool(opt.is_stuck[()]) and ref.stuck
    assert_bitwise(opt.current_point, ref.point, "converged point")
    assert np.abs(opt.current_gradient).max() < 1e-5
    with pytest.raises(AssertionError):
        dz.LBFGSOptimizer(None, EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 0.0, 8)     # @assert step > 0


# ----------------------------------------------------------------------------- live AdGDOptimizer (:179-312)
@pytest.mark.parametrize("n,tree", [(2, False), (10, True), (64, True)])
def test_adgd_c_oracle_equals_python_restatement(orc, n, tree):
    import dzo_oracle_py as P
    x0 = _x0(orc, n, 7) * 0.5
    py = P.LiveAdGDOptimizer(P.Rosenbrock(tree), list(x0), 1e-3, tree)
    c = orc.AdGD(ROSEN, x0, 1e-3, orc.TREE if tree else orc.SEQ)
    for it in range(60):
        py.step(); c.step(1)
        assert_bitwise(c.point, np.array(py.current_point), f"iter {it} point")
        assert_bitwise(c.delta_gradient, np.array(py.delta_gradient), f"iter {it} delta_gradient")
        s = c.scalars
        assert s[0] == py.current_objective_value and s[1] == py.delta_objective_value
        assert s[2] == py.current_step_size and s[3] == py.previous_step_size
        assert int(s[4]) == py.iteration_count and bool(s[5]) == py.is_stuck


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 34, 4096, 20000])
def test_gpu_adgd_trace(gpu, orc, n):
    dz = gpu
    EF = dz.ExampleFunctions
    x0 = _x0(orc, n, 8) * 0.5
    opt = dz.AdGDOptimizer(None, EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1e-3)
    ref = orc.AdGD(ROSEN, x0, 1e-3, orc.TREE)

    def compare(tag):
        assert_bitwise(opt.current_point, ref.point, f"{tag}: point")
        assert_bitwise(opt.delta_point, ref.delta_point, f"{tag}: delta_point")
        assert_bitwise(opt.current_gradient, ref.gradient, f"{tag}: gradient")
        assert_bitwise(opt.delta_gradient, ref.delta_gradient, f"{tag}: delta_gradient")
        assert_bitwise(opt._scalars(), ref.scalars, f"{tag}: scalars (f, df, step sizes, iteration, stuck)")

    compare("ctor")
    for it in range(30):
        dz.step_(opt); ref.step(1)
        compare(f"n={n} iter {it}")
    opt.step(200); ref.step(200)
    compare("fused")
    assert float(opt.current_objective_value[()]) < orc.objective(ROSEN, x0, orc.TREE)[0]


# ----------------------------------------------------------------------------- live LineSearchEvaluator (:12-92)
@pytest.mark.parametrize("n,tree", [(2, False), (10, True), (64, True)])
def test_line_search_evaluator_c_oracle_equals_python_restatement(orc, n, tree):
    import dzo_oracle_py as P
    x = _x0(orc, n, 9)
    fn = P.Rosenbrock(tree)
    g = [0.0] * n
    fn.g(g, list(x))
    d = -np.array(g)
    order = orc.TREE if tree else orc.SEQ
    f_old = fn.f(list(x))
    overlap = P.dot(g, list(d), tree)
    for step, cg in ((1e-3, True), (1e-4, False), (0.5, True)):
        tp, tg, res = orc.line_search_evaluate(ROSEN, x, f_old, d, overlap, step, cg, order)
        ptp, ptg, pf, pir, psr = P.live_line_search_evaluate(fn, list(x), f_old, list(d), overlap, step, cg, tree)
        assert_bitwise(tp, np.array(ptp), "trial point")
        assert res[0] == pf and res[1] == pir
        if cg:
            assert_bitwise(tg, np.array(ptg), "trial gradient")
            assert res[2] == psr
    # a small step along -g decreases f: Armijo ratio in (0, 1], curvature ratio below 1
    tp, tg, res = orc.line_search_evaluate(ROSEN, x, f_old, d, overlap, 1e-6, True, order)
    assert 0.0 < res[1] <= 1.0 + 1e-6 and res[2] < 1.0 + 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 34, 4096, 20000])
def test_gpu_line_search_evaluator(gpu, orc, n):
    dz = gpu
    EF = dz.ExampleFunctions
    x = _x0(orc, n, 9)
    g = orc.gradient(ROSEN, x[None, :], orc.TREE)[0]
    f_old = orc.objective(ROSEN, x[None, :], orc.TREE)[0]
    d = -g
    overlap = float(orc.dot(g, d, orc.TREE))
    lse = dz.LineSearchEvaluator(None, EF.rosenbrock_function, EF.rosenbrock_gradient_, x, f_old, g, d, overlap)
    for step, cg in ((1e-3, True), (1e-4, False), (0.5, True), (0.0, True)):
        f_new = lse(step, cg)
        tp, tg, res = orc.line_search_evaluate(ROSEN, x, f_old, d, overlap, step, cg, orc.TREE)
        assert_bitwise(lse.trial_point, tp, f"step {step}: trial point")
        assert_bitwise([f_new, float(lse.improvement_ratio[()])], res[:2], f"step {step}: f_new, improvement_ratio")
        if cg:
            assert_bitwise(lse.trial_gradient, tg, f"step {step}: trial gradient")
            assert_bitwise(float(lse.slope_ratio[()]), res[2], f"step {step}: slope_ratio")


# ----------------------------------------------------------------------------- DZO_ORDER_TREE_BLOCKED (n > 65536)
def test_blocked_order_is_the_tree_order_up_to_one_block(orc):
    import ctypes as C
    x, y = _x0(orc, 70000, 11), _x0(orc, 70000, 12)
    out = C.c_double()
    for n in (2, 4097, 65536):
        orc.lib().dzo_cpu_dot(orc.TREE_BLOCKED, n, orc._dp(x), orc._dp(y), C.byref(out))
        assert out.value == float(orc.dot(x[:n], y[:n], orc.TREE))
        assert orc.objective(ROSEN, x[None, :n - n % 2], orc.TREE_BLOCKED)[0] == orc.objective(ROSEN, x[None, :n - n % 2], orc.TREE)[0]
    # above one block: block results added in ascending order (checked against a numpy-free restatement)
    import dzo_oracle_py as P
    orc.lib().dzo_cpu_dot(orc.TREE_BLOCKED, 70000, orc._dp(x), orc._dp(y), C.byref(out))
    assert out.value == P.dot(list(x), list(y), P.BLOCKED)
    b0 = float(orc.dot(x[:65536], y[:65536], orc.TREE)); b1 = float(orc.dot(x[65536:], y[65536:], orc.TREE))
    assert out.value == b0 + b1
    assert orc.objective(ROSEN, x[None, :], orc.TREE_BLOCKED)[0] == P.Rosenbrock(P.BLOCKED).f(list(x))


def test_lbfgs_blocked_c_oracle_equals_python_restatement(orc):
    import dzo_oracle_py as P
    n = 65536 + 40
    x0 = _x0(orc, n, 3)
    py = P.LiveLBFGSOptimizer(P.Rosenbrock(P.BLOCKED), list(x0), 0.5, 2, P.BLOCKED)
    c = orc.LBFGS(ROSEN, x0, 0.5, 2, orc.TREE_BLOCKED)
    for it in range(3):
        py.step(); c.step(1)
        assert_bitwise(c.point, np.array(py.current_point), f"iter {it} point")
        assert_bitwise(c.direction, np.array(py.step_direction), f"iter {it} direction")
        assert c.objective == py.current_objective_value
        assert_bitwise(c.rho_history, np.array(py.rho), "rho history")


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(65538, 2), (131072, 3), (200000, 5), (1 << 20, 10), (1 << 21, 4), (3_000_000, 2)])
def test_gpu_lbfgs_grid_wide_trace(gpu, orc, n, m):
    """n > DZO_TREE_BLOCK: cooperative grid, one cluster per block, DZO_ORDER_TREE_BLOCKED; bitwise vs the oracle."""
    import ctypes as C
    dz = gpu
    EF = dz.ExampleFunctions
    x0 = _x0(orc, n, 5)
    opt = dz.LBFGSOptimizer(None, EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, m)
    order, clusters = C.c_int(), C.c_int()
    assert dz.lib().dzo_lbfgs_info(opt._h, None, C.byref(order), C.byref(clusters)) == 0
    assert order.value == 2 and 8 < clusters.value <= 8 * ((n + 65535) // 65536)      # CTAs of the cooperative grid
    ref = orc.LBFGS(ROSEN, x0, 1.0, m, orc.TREE_BLOCKED)

    def compare(tag):
        assert_bitwise(opt.current_point, ref.point, f"{tag}: point")
        assert_bitwise(opt.delta_point, ref.delta_point, f"{tag}: delta_point")
        assert_bitwise(opt.current_gradient, ref.gradient, f"{tag}: gradient")
        assert_bitwise(opt.delta_gradient, ref.delta_gradient, f"{tag}: delta_gradient")
        assert_bitwise(opt.step_direction, ref.direction, f"{tag}: direction")
        assert float(opt.current_objective_value[()]) == ref.objective
        assert float(opt.delta_objective_value[()]) == ref.delta_objective
        assert int(opt.iteration_count[()]) == ref.iteration_count and bool(opt.is_stuck[()]) == ref.stuck
        assert_bitwise(opt.rho_history, ref.rho_history, f"{tag}: rho")

    compare("ctor")
    steps = 4 if n >= (1 << 20) else 12
    for it in range(steps):
        dz.step_(opt); ref.step(1)
        compare(f"n={n} iter {it}")
    opt.step(15); ref.step(15)            # 15 step! calls in one cooperative launch
    compare("fused")


@pytest.mark.gpu
@pytest.mark.parametrize("n", [65538, 200000, 1 << 20])
def test_gpu_adgd_grid_wide_trace(gpu, orc, n):
    """AdGD above n = 65536: cooperative grid, DZO_ORDER_TREE_BLOCKED; bitwise vs the oracle."""
    dz = gpu
    EF = dz.ExampleFunctions
    x0 = _x0(orc, n, 8) * 0.5
    opt = dz.AdGDOptimizer(None, EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1e-3)
    ref = orc.AdGD(ROSEN, x0, 1e-3, orc.TREE_BLOCKED)

    def compare(tag):
        assert_bitwise(opt.current_point, ref.point, f"{tag}: point")
        assert_bitwise(opt.delta_point, ref.delta_point, f"{tag}: delta_point")
        assert_bitwise(opt.current_gradient, ref.gradient, f"{tag}: gradient")
        assert_bitwise(opt.delta_gradient, ref.delta_gradient, f"{tag}: delta_gradient")
        assert_bitwise(opt._scalars(), ref.scalars, f"{tag}: scalars")

    compare("ctor")
    for it in range(8):
        dz.step_(opt); ref.step(1)
        compare(f"n={n} iter {it}")
    opt.step(40); ref.step(40)
    compare("fused")


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["recycle", "barrier", "no-stage"])
def test_gpu_grid_reduction_variants(gpu, orc, variant):
    """The grid-wide kernels reduce through flagged 16-byte lines in per-CTA inboxes (grid_lbfgs.cuh).  (recycle) a handle
    whose line tags start just below the recycling threshold 2^27 crosses it between launches -- every CTA clears its
    inbox and the tags start over at 1; (barrier) the grid.sync() reductions kept for A/B; (no-stage) passes read from
    global memory instead of the staged copies.  All three walk the oracle's trajectory bit for bit."""
    dz = gpu
    EF = dz.ExampleFunctions
    n, m = 200000, 4
    x0 = _x0(orc, n, 11)
    knobs = {"recycle": ("grid_ll_first_seq", (1 << 27) - 40, 1), "barrier": ("grid_ll", 0, 1), "no-stage": ("grid_stage", 0, 1)}
    key, value, default = knobs[variant]
    try:
        dz.set_tuning(key, value)
        opt = dz.LBFGSOptimizer(None, EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, m)
        leg = dz.LegacyLBFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, dz.QuadraticLineSearch(0), x0, 1.0, m)
    finally:
        dz.set_tuning(key, default)
    ref = orc.LBFGS(ROSEN, x0, 1.0, m, orc.TREE_BLOCKED)
    for it in range(10):                                   # ~10 reductions per launch: the threshold falls inside this loop
        dz.step_(opt); ref.step(1)
        assert_bitwise(opt.current_point, ref.point, f"{variant} iter {it}: point")
        assert_bitwise(opt.step_direction, ref.direction, f"{variant} iter {it}: direction")
        assert float(opt.current_objective_value[()]) == ref.objective
    opt.step(20); ref.step(20)
    assert_bitwise(opt.current_point, ref.point, f"{variant}: fused steps")
    # the legacy kernel shares the reduction code: compare two handles created with and without the knob
    leg2 = dz.LegacyLBFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, dz.QuadraticLineSearch(0), x0, 1.0, m)
    for _ in range(6):
        dz.step_(leg); dz.step_(leg2)
    leg.step(6); leg2.step(6)
    assert_bitwise(leg.current_point, leg2.current_point, f"{variant}: legacy L-BFGS")
