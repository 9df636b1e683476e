"""Legacy LBFGSOptimizer (legacy/DZOptimization.jl:458-695) and the L2 / uniform-box decorators (:222-296):
oracle vs the independent Python restatement (CPU) and CUDA vs oracle (GPU), all bitwise."""
import numpy as np
import pytest

from conftest import assert_bitwise

ROSEN = 1

# (l2_lambda, box) decorations exercised everywhere
DECOR = [(None, None), (0.1, None), (None, (-0.5, 0.8)), (0.01, (-1.0, 0.5))]


def _x0(orc, n, seed):
    return 4.0 * orc.pcg_fill(n, seed) - 2.0


def _py_fn(P, tree, lam, box):
    fn = P.Rosenbrock(tree)
    if lam is not None:
        fn = P.L2Regularized(fn, lam)
    if box is not None:
        fn = P.UniformBox(fn, *box)
    return fn


@pytest.mark.parametrize("n,m,tree,mi", [(2, 3, False, 0), (10, 4, True, 0), (16, 1, True, 2), (34, 5, True, 0)])
@pytest.mark.parametrize("lam,box", DECOR)
def test_c_oracle_equals_python_restatement(orc, n, m, tree, mi, lam, box):
    import dzo_oracle_py as P
    x0 = _x0(orc, n, 5)
    py = P.LegacyLBFGSOptimizer(_py_fn(P, tree, lam, box), list(x0), 1.0, m, mi, tree)
    c = orc.LegacyLBFGS(ROSEN, x0, 1.0, m, mi, lam, box, orc.TREE if tree else orc.SEQ)
    for it in range(30):
        py.step(); c.step(1)
        assert_bitwise(c.point, np.array(py.current_point), f"iter {it} point")
        assert_bitwise(c.delta_point, np.array(py.delta_point), f"iter {it} delta_point")
        assert_bitwise(c.gradient, np.array(py.current_gradient), f"iter {it} gradient")
        assert_bitwise(c.delta_gradient, np.array(py.delta_gradient), f"iter {it} delta_gradient")
        assert_bitwise(c.direction, np.array(py.next_step_direction), f"iter {it} direction")
        s = c.scalars
        assert s[0] == py.current_objective_value and s[1] == py.delta_objective_value and s[2] == py.last_step_length
        assert int(s[3]) == py.iteration_count and bool(s[4]) == py.has_terminated and int(s[5]) == py._history_count
        rho, alpha = c.history
        assert_bitwise(rho, np.array(py._rho), "rho")
        assert_bitwise(alpha, np.array(py._alpha), "alpha")


def test_legacy_lbfgs_converges_to_the_rosenbrock_minimum(orc):
    c = orc.LegacyLBFGS(ROSEN, _x0(orc, 2, 5), 1.0, 3, order=orc.SEQ)
    c.step(500)
    s = c.scalars
    assert bool(s[4]) and np.abs(c.point - 1.0).max() < 1e-7 and s[0] < 1e-14
    # zero gradient at construction => has_terminated (:524-526)
    z = orc.LegacyLBFGS(ROSEN, np.ones(4), 1.0, 3)
    assert bool(z.scalars[4]) and int(z.scalars[3]) == 0


def test_decorators_do_what_the_reference_wrappers_do(orc):
    # box: every iterate stays inside the box and the minimiser sits on the active bounds
    c = orc.LegacyLBFGS(ROSEN, _x0(orc, 16, 5), 1.0, 5, box=(-0.5, 0.8), order=orc.SEQ)
    for _ in range(200):
        c.step(1)
        x = c.point
        assert x.min() >= -0.5 and x.max() <= 0.8
    g = c.gradient
    assert np.all(g[(c.point <= -0.5)] <= 0.0) and np.all(g[(c.point >= 0.8)] >= 0.0)   # masked entries are 0
    # L2: the regularised objective is what the optimizer reports
    lam = 0.1
    d = orc.LegacyLBFGS(ROSEN, _x0(orc, 16, 5), 1.0, 2, l2_lambda=lam, order=orc.SEQ)
    d.step(30)
    x = d.point
    f = orc.objective(ROSEN, x, orc.SEQ)[0] + lam * float(np.sum(x * x))
    assert abs(d.scalars[0] - f) <= 1e-12 * abs(f)


def test_abi_rejects_bad_legacy_arguments(orc):
    import ctypes as C
    x = np.ones(4)
    h = C.c_void_p()
    p = x.ctypes.data_as(C.POINTER(C.c_double))
    L = orc.lib()
    assert L.dzo_cpu_legacy_lbfgs_create(C.byref(h), ROSEN, 0, 0, 4, p, 1.0, 0, 0, 0, 0.0, 0.0, 0.0, orc.TREE) == -1   # history_length > 0
    assert L.dzo_cpu_legacy_lbfgs_create(C.byref(h), ROSEN, 0, 0, 4, p, 1.0, 3, 0, 8, 0.0, 0.0, 0.0, orc.TREE) == -1   # unknown decorator bit
    assert L.dzo_cpu_legacy_lbfgs_create(C.byref(h), ROSEN, 0, 0, 4, p, 1.0, 3, 0, 2, 0.0, 1.0, -1.0, orc.TREE) == -1  # lower > upper
    assert L.dzo_cpu_legacy_lbfgs_create(C.byref(h), 2, 0, 2, 4, p, 1.0, 3, 0, 1, 0.1, 0.0, 0.0, orc.TREE) == -5       # decorators: Rosenbrock only


def test_host_mirror_peels_wrappers(dz):
    EF = dz.ExampleFunctions
    f, g = EF.rosenbrock_function, EF.rosenbrock_gradient_
    r = dz._resolve_decorated(dz.UniformBoxConstraint(-1, 2), dz.L2RegularizationWrapper(f, 0.5),
                              dz.UniformBoxGradientWrapper(dz.L2GradientWrapper(g, 0.5), -1, 2))
    assert r == (1, 0, 0.5, (-1.0, 2.0))
    assert dz._resolve_decorated(None, f, g) == (1, 0, None, None)
    with pytest.raises(TypeError):                      # lambda on the objective only
        dz._resolve_decorated(None, dz.L2RegularizationWrapper(f, 0.5), g)
    with pytest.raises(TypeError):                      # box on the gradient only
        dz._resolve_decorated(None, f, dz.UniformBoxGradientWrapper(g, 0, 1))
    with pytest.raises(TypeError):                      # wrong nesting order
        dz._resolve_decorated(dz.UniformBoxConstraint(0, 1), dz.L2RegularizationWrapper(f, 0.5),
                              dz.L2GradientWrapper(dz.UniformBoxGradientWrapper(g, 0, 1), 0.5))
    with pytest.raises(TypeError):
        dz.LegacyLBFGSOptimizer(f, g, None, np.ones(4), 1.0, 3)      # linesearch must be a QuadraticLineSearch


def _gpu_pair(dz, orc, x0, m, mi, lam, box):
    EF = dz.ExampleFunctions
    f, g, c = EF.rosenbrock_function, EF.rosenbrock_gradient_, dz.NULL_CONSTRAINT
    if lam is not None:
        f, g = dz.L2RegularizationWrapper(f, lam), dz.L2GradientWrapper(g, lam)
    if box is not None:
        g, c = dz.UniformBoxGradientWrapper(g, *box), dz.UniformBoxConstraint(*box)
    opt = dz.LegacyLBFGSOptimizer(c, f, g, dz.QuadraticLineSearch(mi), x0, 1.0, m)
    ref = orc.LegacyLBFGS(ROSEN, x0, 1.0, m, mi, lam, box, orc.TREE)
    return opt, ref


def _compare(opt, ref, tag):
    assert_bitwise(opt.current_point, ref.point, f"{tag}: point")
    assert_bitwise(opt.delta_point, ref.delta_point, f"{tag}: delta_point")
    assert_bitwise(opt.current_gradient, ref.gradient, f"{tag}: gradient")
    assert_bitwise(opt.delta_gradient, ref.delta_gradient, f"{tag}: delta_gradient")
    assert_bitwise(opt.next_step_direction, ref.direction, f"{tag}: direction")
    assert_bitwise(opt._scalars(), ref.scalars, f"{tag}: scalars (f, df, L, iteration, terminated, history count)")
    rho, alpha = ref.history
    assert_bitwise(opt._rho, rho, f"{tag}: rho")
    assert_bitwise(opt._alpha, alpha, f"{tag}: alpha")


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,mi", [(2, 2, 0), (34, 5, 0), (2048, 8, 3), (16384, 10, 0), (20000, 3, 0)])
@pytest.mark.parametrize("lam,box", DECOR)
def test_gpu_legacy_lbfgs_trace(gpu, orc, n, m, mi, lam, box):
    opt, ref = _gpu_pair(gpu, orc, _x0(orc, n, 5), m, mi, lam, box)
    _compare(opt, ref, "ctor")
    for it in range(20):
        gpu.step_(opt); ref.step(1)
        _compare(opt, ref, f"n={n} iter {it}")
    opt.step(25); ref.step(25)            # 25 step! calls in one cluster-kernel launch
    _compare(opt, ref, "fused")


@pytest.mark.gpu
@pytest.mark.parametrize("lam,box", DECOR)
def test_gpu_legacy_lbfgs_to_termination(gpu, orc, lam, box):
    opt, ref = _gpu_pair(gpu, orc, _x0(orc, 16, 5), 5, 0, lam, box)
    for _ in range(100):
        opt.step(100); ref.step(100)
        if opt.has_terminated[()]:
            break
    assert bool(opt.has_terminated[()]) and bool(ref.scalars[4])
    _compare(opt, ref, "terminated")
    if box is not None:
        x = opt.current_point
        assert x.min() >= box[0] and x.max() <= box[1]


@pytest.mark.gpu
def test_gpu_legacy_lbfgs_degenerate_starts(gpu, orc):
    # zero gradient at construction (terminated at once), a start on the box boundary, a huge start
    for x0, box in ((np.ones(8), None), (np.full(8, 0.8), (-0.5, 0.8)), (np.full(8, 1e6), None), (np.full(8, 3.0), (-1.0, 1.0))):
        opt, ref = _gpu_pair(gpu, orc, x0, 4, 0, None, box)
        _compare(opt, ref, "ctor")
        for it in range(30):
            gpu.step_(opt); ref.step(1)
            _compare(opt, ref, f"iter {it}")


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,mi,lam,box", [(65538, 2, 0, None, None), (131072, 3, 2, 0.1, None), (200000, 5, 0, None, (-0.5, 0.8)),
                                            (1 << 20, 10, 0, 0.01, (-1.0, 0.5)), (3_000_000, 2, 0, None, None)])
def test_gpu_legacy_lbfgs_grid_wide_trace(gpu, orc, n, m, mi, lam, box):
    """n > DZO_TREE_BLOCK: cooperative grid, eight CTAs per block, DZO_ORDER_TREE_BLOCKED; bitwise vs the oracle."""
    x0 = _x0(orc, n, 5)
    opt, _unused = _gpu_pair(gpu, orc, x0[:4], 1, 0, None, None)      # (exercise the small path once more)
    opt.close()
    EF = gpu.ExampleFunctions
    f, g, c = EF.rosenbrock_function, EF.rosenbrock_gradient_, gpu.NULL_CONSTRAINT
    if lam is not None:
        f, g = gpu.L2RegularizationWrapper(f, lam), gpu.L2GradientWrapper(g, lam)
    if box is not None:
        g, c = gpu.UniformBoxGradientWrapper(g, *box), gpu.UniformBoxConstraint(*box)
    opt = gpu.LegacyLBFGSOptimizer(c, f, g, gpu.QuadraticLineSearch(mi), x0, 1.0, m)
    ref = orc.LegacyLBFGS(ROSEN, x0, 1.0, m, mi, lam, box, orc.TREE_BLOCKED)
    _compare(opt, ref, "ctor")
    for it in range(4 if n >= (1 << 20) else 10):
        gpu.step_(opt); ref.step(1)
        _compare(opt, ref, f"n={n} iter {it}")
    opt.step(12); ref.step(12)
    _compare(opt, ref, "fused")
