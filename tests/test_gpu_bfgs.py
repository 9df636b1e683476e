"""GPU parity tests of the BFGS step! path: CUDA (through the C ABI) vs the CPU oracle.

Bar (BASELINE.json north_star): per-iteration iterates, objective values and step sizes within
1e-12 relative for the first k iterations, converged points within 1e-9.  Because the CUDA
kernels and the oracle share one summation order and never contract to FMA, the tests assert
the stronger property: BITWISE equality of every field after every step!.
"""
import numpy as np
import pytest

from conftest import assert_bitwise

pytestmark = pytest.mark.gpu

SEQ, TREE = 0, 1
ROSEN = 1


def _x0(orc, count, seed):
    return 4.0 * orc.pcg_fill(count, seed) - 2.0


def _sym(orc, n, seed):
    """random bitwise-symmetric matrix (the inverse Hessian is always bitwise symmetric)"""
    a = orc.pcg_fill(n * n, seed).reshape(n, n)
    return (a + a.T) + n * np.eye(n)


# ----------------------------------------------------------------------------- kernel-level rows of SURVEY 8a
@pytest.mark.parametrize("n", [2, 5, 16, 32])
def test_dot_sequential(gpu, orc, n):
    import dev
    v, w = _x0(orc, n, 11), _x0(orc, n, 12)
    assert dev.dot(v, w, SEQ) == orc.dot(v, w, orc.SEQ)


@pytest.mark.parametrize("n", [1, 2, 31, 4096, 8191, 8192, 16384, 100003])
def test_dot_tree(gpu, orc, n):
    import dev
    v, w = _x0(orc, n, 13), _x0(orc, n, 14)
    assert dev.dot(v, w, TREE) == orc.dot(v, w, orc.TREE)


@pytest.mark.parametrize("n", [2, 7, 16, 32])
def test_gemv_sequential(gpu, orc, n):
    import dev
    H, v = _sym(orc, n, 21), _x0(orc, n, 22)
    assert_bitwise(dev.gemv(H, v, SEQ), orc.gemv(H, v, orc.SEQ), "gemv seq")


@pytest.mark.parametrize("n", [33, 64, 255, 256, 1000, 1024, 1025, 2050, 3073])
def test_gemv_tree(gpu, orc, n):
    import dev
    a = orc.pcg_fill(n * n, 23).reshape(n, n) - 0.5      # NOT symmetric: pins the row/column convention
    v = _x0(orc, n, 24)
    assert_bitwise(dev.gemv(a, v, TREE), orc.gemv(a, v, orc.TREE), "gemv tree")


@pytest.mark.parametrize("n", [2, 16, 32])
def test_update_inverse_hessian_sequential(gpu, orc, n):
    import dev
    H, d, dg, g = _sym(orc, n, 31), _x0(orc, n, 32), _x0(orc, n, 33), _x0(orc, n, 34)
    got = dev.update_inverse_hessian(H, -0.37, d, dg, g, SEQ)
    ref = orc.update_inverse_hessian(H, -0.37, d, dg, g, orc.SEQ)
    for a, b, what in zip(got, ref, ("H", "step_direction", "scratch", "next_direction")):
        assert_bitwise(a, b, what)


@pytest.mark.parametrize("n,fused", [(34, True), (256, True), (1000, False), (1026, True), (2051, True), (3072, True)])
def test_update_inverse_hessian_tree(gpu, orc, n, fused):
    import dev
    H, d, dg, g = _sym(orc, n, 35), _x0(orc, n, 36), _x0(orc, n, 37), _x0(orc, n, 38)
    got = dev.update_inverse_hessian(H, -0.37, d, dg, g if fused else None, TREE)
    ref = orc.update_inverse_hessian(H, -0.37, d, dg, g if fused else None, orc.TREE)
    for a, b, what in zip(got, ref, ("H", "step_direction", "scratch", "next_direction")):
        if a is None and b is None:
            continue
        assert_bitwise(a, b, what)
    # the update keeps H bitwise symmetric
    assert_bitwise(got[0], got[0].T, "symmetry")


@pytest.mark.parametrize("n", [1, 2, 255, 513, 2048])
def test_identity(gpu, n):
    import dev
    assert_bitwise(dev.identity(n), np.eye(n), "identity")


@pytest.mark.parametrize("n,order", [(2, SEQ), (16, SEQ), (32, SEQ), (34, TREE), (2048, TREE), (20000, TREE)])
def test_rosenbrock_objective_gradient(gpu, orc, n, order):
    import dev
    x = _x0(orc, 5 * n, 41).reshape(5, n)
    assert_bitwise(dev.objective(ROSEN, x, order), orc.objective(ROSEN, x, order), "objective")
    assert_bitwise(dev.gradient(ROSEN, x, order), orc.gradient(ROSEN, x, order), "gradient")


@pytest.mark.parametrize("n,order", [(2, SEQ), (16, SEQ), (64, TREE), (4096, TREE), (16384, TREE)])
@pytest.mark.parametrize("t1", [1.0, 1e-3, 37.0, 1e-300, 1e300])
def test_line_search(gpu, orc, n, order, t1):
    import dev
    x = _x0(orc, n, 51)
    g = orc.gradient(ROSEN, x, order)[0]
    f0 = orc.objective(ROSEN, x, order)[0]
    got = dev.line_search(ROSEN, x, g, f0, t1, order)
    ref = orc.line_search(ROSEN, x, g, f0, t1, order)
    assert got == ref


def test_line_search_degenerate(gpu, orc):
    import dev
    x = _x0(orc, 16, 52)
    f0 = orc.objective(ROSEN, x, SEQ)[0]
    z = np.zeros(16)
    for order, xx in ((SEQ, x), (TREE, np.tile(x, 4))):
        ff = orc.objective(ROSEN, xx, order)[0]
        zz = np.zeros(xx.size)
        assert dev.line_search(ROSEN, xx, zz, ff, 1.0, order) == (0.0, ff)            # zero direction :71-85
        assert dev.line_search(ROSEN, xx, xx, ff, float("inf"), order) == (0.0, ff)   # t1 = L/0
        assert dev.line_search(ROSEN, xx, xx, float("inf"), 1.0, order)[0] == 0.0     # non-finite f0 :64-66
    assert f0 > 0 and z.sum() == 0


# ----------------------------------------------------------------------------- step! traces
FIELDS = ("point", "gradient", "delta_point", "delta_gradient", "direction", "objective", "step_length")


def _compare_state(opt, ref, batched, tag):
    get = {
        "point": opt.current_point, "gradient": opt.current_gradient, "delta_point": opt.delta_point,
        "delta_gradient": opt.delta_gradient, "direction": opt.next_step_direction,
        "objective": opt.current_objective_value, "step_length": opt.last_step_length,
    }
    for name in FIELDS:
        r = getattr(ref, name)
        if not batched:
            r = r[0]
        assert_bitwise(np.asarray(get[name]), np.asarray(r), f"{tag}: {name}")
    rt = ref.step_type if batched else ref.step_type[0]
    ri = ref.iteration_count if batched else ref.iteration_count[0]
    rz = ref.terminated if batched else ref.terminated[0]
    assert np.array_equal(np.asarray(opt.last_step_type), np.asarray(rt)), f"{tag}: step type"
    assert np.array_equal(np.asarray(opt.iteration_count), np.asarray(ri)), f"{tag}: iteration count"
    assert np.array_equal(np.asarray(opt.has_converged), np.asarray(rz)), f"{tag}: has_converged"


def test_readme_rosenbrock_n2(gpu, orc):
    """config 1: README Rosenbrock n=2 from rand(2) (PCG seeds 0..999), step size 1.0, to has_converged."""
    dz = gpu
    EF = dz.ExampleFunctions
    x0 = np.stack([orc.pcg_fill(2, s) for s in range(1000)])
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
    ref = orc.BFGS(ROSEN, x0, 1.0, order=orc.SEQ, nthreads=8)
    _compare_state(opt, ref, True, "ctor")
    for it in range(12):
        dz.step_(opt); ref.step(1)
        _compare_state(opt, ref, True, f"iter {it}")
    for _ in range(40):
        opt.step(25); ref.step(25)
        if opt.count_active() == 0:
            break
    _compare_state(opt, ref, True, "converged")
    assert opt.count_active() == 0 and ref.count_active() == 0
    assert opt.has_converged.all()
    assert np.abs(opt.current_point - 1.0).max() < 1e-6          # analytic minimum (1, 1)
    for p in (0, 17, 999):
        assert_bitwise(opt.inverse_hessian(p), ref.inverse_hessian(p), "H")


@pytest.mark.parametrize("n,batch", [(2, 517), (4, 300), (6, 777), (8, 64), (10, 333), (12, 1000), (14, 95), (16, 2000), (18, 130),
                                     (20, 257), (22, 33), (24, 500), (26, 31), (28, 200), (30, 70), (32, 129)])
def test_batched_trace(gpu, orc, n, batch):
    dz = gpu
    EF = dz.ExampleFunctions
    x0 = _x0(orc, n * batch, 2024).reshape(batch, n)
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
    ref = orc.BFGS(ROSEN, x0, 1.0, order=orc.SEQ, nthreads=8)
    assert opt.summation_order == SEQ
    for it in range(25):
        dz.step_(opt); ref.step(1)
        _compare_state(opt, ref, True, f"n={n} iter {it}")
    opt.step(40); ref.step(40)                                   # k fused steps in ONE launch == 40 step! calls
    _compare_state(opt, ref, True, f"n={n} after fused steps")
    for p in (0, batch - 1):
        assert_bitwise(opt.inverse_hessian(p), ref.inverse_hessian(p), "H")
        assert_bitwise(opt.inverse_hessian(p), opt.inverse_hessian(p).T, "H symmetric")


def test_batched_to_convergence(gpu, orc):
    dz = gpu
    EF = dz.ExampleFunctions
    n, batch = 16, 512
    x0 = _x0(orc, n * batch, 7).reshape(batch, n)
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
    ref = orc.BFGS(ROSEN, x0, 1.0, order=orc.SEQ, nthreads=8)
    for _ in range(200):
        opt.step(50); ref.step(50)
        if opt.count_active() == 0:
            break
    assert opt.count_active() == 0 and ref.count_active() == 0
    _compare_state(opt, ref, True, "converged")
    # every problem sits at a stationary point; the global minimiser is (1,...,1), f = 0
    at_global = np.abs(opt.current_point - 1.0).max(axis=1) < 1e-6
    assert at_global.mean() > 0.5
    assert (opt.current_objective_value[at_global] < 1e-10).all()


@pytest.mark.parametrize("n", [34, 64, 130, 500, 512, 514, 2048, 4098, 16384])
def test_large_trace(gpu, orc, n):
    dz = gpu
    EF = dz.ExampleFunctions
    x0 = _x0(orc, n, 1)
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0)
    ref = orc.BFGS(ROSEN, x0[None, :], 1.0, order=orc.TREE, nthreads=8)
    assert opt.summation_order == TREE
    _compare_state(opt, ref, False, "ctor")
    steps = 12 if n < 16384 else 6
    types = []
    for it in range(steps):
        dz.step_(opt); ref.step(1)
        _compare_state(opt, ref, False, f"n={n} iter {it}")
        types.append(int(opt.last_step_type[()]))
    assert dz.StepType.BFGSStep in types                         # the n^2 kernels really ran
    if n <= 4098:
        H = opt.inverse_hessian()
        assert_bitwise(H, ref.inverse_hessian(0), "H")
        assert_bitwise(H, H.T, "H symmetric")


def test_large_to_convergence(gpu, orc):
    dz = gpu
    EF = dz.ExampleFunctions
    n = 64
    x0 = _x0(orc, n, 5)
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0)
    ref = orc.BFGS(ROSEN, x0[None, :], 1.0, order=orc.TREE)
    for _ in range(100):
        opt.step(50); ref.step(50)
        if opt.has_converged[()]:
            break
    assert bool(opt.has_converged[()]) and bool(ref.terminated[0])
    _compare_state(opt, ref, False, "converged")
    # converged point within 1e-9 of the oracle's (it is bitwise equal) and stationary
    assert np.abs(opt.current_point - ref.point[0]).max() <= 1e-9
    assert np.abs(opt.current_gradient).max() < 1e-5


def test_set_state_resume(gpu, orc):
    """save/load in the middle of optimization (README.md:11; legacy/DZOptimization.jl:819-862)."""
    dz = gpu
    EF = dz.ExampleFunctions
    for n, batched, batch in ((16, True, 50), (2048, False, 1)):
        x0 = _x0(orc, n * batch, 9).reshape(batch, n)
        mk = lambda: dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0 if batched else x0[0], 1.0,
                                      batched=batched)
        a = mk()
        a.step(7)
        if batched:
            H = np.stack([a.inverse_hessian(p) for p in range(batch)])
        else:
            H = a.inverse_hessian()
        saved = (a.current_point, H, a.delta_point, a.delta_gradient, a.last_step_length, a.last_step_type,
                 a.iteration_count)
        b = mk()
        b.set_state(*saved)
        assert_bitwise(b.current_objective_value, a.current_objective_value, "f after resume")
        assert_bitwise(b.current_gradient, a.current_gradient, "g after resume")
        assert_bitwise(b.next_step_direction, a.next_step_direction, "d after resume")
        # the oracle's own state-rebuilding constructor (:819-862), fed the oracle's own saved fields
        order = orc.SEQ if batched else orc.TREE
        ra = orc.BFGS(ROSEN, x0, 1.0, order=order)
        ra.step(7)
        Hr = np.stack([ra.inverse_hessian(p) for p in range(batch)])
        for got, want, what in ((saved[0], ra.point, "point"), (H.reshape(Hr.shape), Hr, "H"), (saved[2], ra.delta_point, "dx"),
                                (saved[3], ra.delta_gradient, "dg"), (saved[4], ra.step_length, "L")):
            assert_bitwise(np.asarray(got).reshape(np.asarray(want).shape), want, f"saved {what} vs oracle")
        rb = orc.BFGS(ROSEN, x0, 1.0, order=order)
        rb.set_state(ra.point, Hr, ra.delta_point, ra.delta_gradient, ra.step_length, ra.step_type, ra.iteration_count)
        _compare_state(b, rb, batched, "after set_state vs the oracle's set_state")
        for it in range(5):
            a.step(1); b.step(1); rb.step(1)
            _compare_state(b, rb, batched, f"resumed iter {it} vs the oracle's resumed optimizer")
        assert_bitwise(b.current_point, a.current_point, "resumed trajectory")
        assert np.array_equal(b.iteration_count, a.iteration_count)
        for p in (0, batch - 1):
            assert_bitwise(b.inverse_hessian(p), rb.inverse_hessian(p), "H after resumed steps")


def test_run_and_test_invariants_full_size(gpu, orc):
    """run_and_test! (legacy/DZOptimization.jl:998-1049) on the device at a size the CPU oracle is not
    asked to follow step by step: 200k problems x n=16.  Recomputed f / g equal the stored fields bitwise,
    delta_point == x_{k+1} - x_k, terminated problems do not move."""
    import dev
    dz = gpu
    EF = dz.ExampleFunctions
    n, batch = 16, 200_000
    x0 = _x0(orc, n * batch, 2024).reshape(batch, n)
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
    prev_x, prev_g = opt.current_point, opt.current_gradient
    for it in range(3):
        was_done = opt.has_converged
        opt.step(1)
        x, g = opt.current_point, opt.current_gradient
        moved = opt.iteration_count > it
        assert_bitwise(dev.objective(ROSEN, x, SEQ), opt.current_objective_value, "f(current_point)")   # :1019-1022
        assert_bitwise(dev.gradient(ROSEN, x, SEQ), g, "g(current_point)")                              # :1025-1032
        dx, dg = opt.delta_point, opt.delta_gradient
        assert_bitwise(dx[moved], (-prev_x[moved]) + x[moved], "delta_point")                           # :1035-1039
        assert_bitwise(dg[moved], (-prev_g[moved]) + g[moved], "delta_gradient")                        # :1042-1046
        assert_bitwise(x[was_done], prev_x[was_done], "terminated problems stay put")
        prev_x, prev_g = x, g
    # spot-check 4096 problems of this big batch against the oracle
    sel = np.arange(0, batch, batch // 4096)[:4096]
    ref = orc.BFGS(ROSEN, x0[sel], 1.0, order=orc.SEQ, nthreads=8)
    ref.step(3)
    assert_bitwise(opt.current_point[sel], ref.point, "spot check")


def test_errors(gpu):
    dz = gpu
    EF = dz.ExampleFunctions
    with pytest.raises(dz.DZOptError) as e:                      # @assert !isnan(f0)  :773
        dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, np.array([np.nan, 1.0]), 1.0)
    assert e.value.code == -3
    with pytest.raises(dz.DZOptError):
        dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, np.zeros(3), 1.0)   # odd n
    with pytest.raises(TypeError):
        dz.BFGSOptimizer(lambda x: 0.0, EF.rosenbrock_gradient_, np.zeros(2), 1.0)           # host closure


# ----------------------------------------------------------------------------- degenerate starts (state, never an error)
@pytest.mark.parametrize("n,batched", [(16, True), (2048, False), (100, False), (300, False)])
def test_degenerate_starts(gpu, orc, n, batched):
    dz = gpu
    EF = dz.ExampleFunctions

    def both(x0, step):
        a = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0 if batched else x0[0], step, batched=batched)
        b = orc.BFGS(ROSEN, x0, step, order=orc.SEQ if batched else orc.TREE)
        return a, b

    # already at the minimiser: gradient 0 -> L/||g|| = Inf -> no probe -> has_converged after one step!, nothing moves
    a, b = both(np.ones((1, n)), 1.0)
    a.step(2); b.step(2)
    _compare_state(a, b, batched, "at minimum")
    assert np.asarray(a.has_converged).all() and int(np.asarray(a.iteration_count).max()) == 0
    # zero initial step length: t1 = 0 -> [GLUE] no probe -> terminates
    a, b = both(np.zeros((1, n)), 0.0)
    a.step(1); b.step(1)
    _compare_state(a, b, batched, "zero step length")
    assert np.asarray(a.has_converged).all()
    # enormous start: objective overflows to +Inf (not NaN) -> constructor accepts (:773), step! terminates (:64-66)
    a, b = both(np.full((1, n), 1e200), 1.0)
    a.step(1); b.step(1)
    _compare_state(a, b, batched, "infinite objective")
    assert np.isinf(np.asarray(a.current_objective_value)).all() and np.asarray(a.has_converged).all()
    # tiny step length: the bracket has to double many times before the point moves (:91-101)
    x0 = _x0(orc, n, 3).reshape(1, n)
    a, b = both(x0, 1e-300)
    for it in range(3):
        a.step(1); b.step(1)
        _compare_state(a, b, batched, f"tiny step iter {it}")


def test_batch_sizes_not_multiple_of_the_warp(gpu, orc):
    dz = gpu
    EF = dz.ExampleFunctions
    for n, batch in ((16, 1), (16, 33), (8, 31), (2, 65), (4, 7)):
        x0 = _x0(orc, n * batch, 77).reshape(batch, n)
        a = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
        b = orc.BFGS(ROSEN, x0, 1.0, order=orc.SEQ)
        a.step(6); b.step(6)
        _compare_state(a, b, True, f"n={n} batch={batch}")


def test_tuning_variants_do_not_change_any_bit(gpu, orc):
    """the A/B knobs (columns in flight, CTA size of the sweeps, line-search kernel, prefetch distance, implicit
    identities) only change scheduling and layout: every variant reproduces the oracle bit for bit."""
    dz = gpu
    EF = dz.ExampleFunctions
    n = 2100
    x0 = _x0(orc, n, 19)
    ref = orc.BFGS(ROSEN, x0[None, :], 1.0, order=orc.TREE)
    ref.step(6)
    for key, values in (("sweep_unroll", (4, 8, 24, 32, 16)), ("sweep_threads", (32, 64, 256, 0)), ("search_variant", (1, 0))):
        for v in values:
            dz.set_tuning(key, v)
            a = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0)
            a.step(6)
            assert_bitwise(a.current_point, ref.point[0], f"{key}={v}: point")
            assert_bitwise(a.next_step_direction, ref.direction[0], f"{key}={v}: direction")
            assert_bitwise(a.inverse_hessian(), ref.inverse_hessian(0), f"{key}={v}: H")
            a.close()
    xb = _x0(orc, 16 * 100, 20).reshape(100, 16)
    rb = orc.BFGS(ROSEN, xb, 1.0)
    rb.step(8)
    for key, values in (("batched_prefetch", (0, 1, 5, 3)), ("batched_lazy", (0, 1))):
        for v in values:
            dz.set_tuning(key, v)
            b = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, xb, 1.0, batched=True)
            b.step(8)
            assert_bitwise(b.current_point, rb.point, f"{key}={v}: batched point")
            assert_bitwise(b.inverse_hessian(99), rb.inverse_hessian(99), f"{key}={v}: batched H")
            for p in range(0, 100, 9):
                assert_bitwise(b.inverse_hessian(p), rb.inverse_hessian(p), f"{key}={v}: batched H[{p}]")
            b.close()


@pytest.mark.parametrize("L0", [1.0, 1e-12, 1e6, 1e-300])
@pytest.mark.parametrize("n", [2, 8, 16, 12])
def test_batched_awkward_inputs(gpu, orc, n, L0):
    """coordinates drawn from a mix of special values (±0, ±1, 1e-8, 1e6, 1e-160, 1e150) and random ones:
    exercises every branch of the per-thread line-search state machine (tiny steps that do not move the
    point, overflowing objectives, zero gradients) against the oracle, bit for bit."""
    dz = gpu
    EF = dz.ExampleFunctions
    batch = 1500
    u = orc.pcg_fill(n * batch, 4242 + n)
    pick = (orc.pcg_fill(n * batch, 777 + n) * 14).astype(int)
    special = np.array([0.0, -0.0, 1.0, -1.0, 1e-8, -1e-8, 1e6, -1e6, 1e-160, 1e150])
    x0 = np.where(pick < 10, special[np.minimum(pick, 9)], 6.0 * u - 3.0).reshape(batch, n)
    x0[0] = 1.0                                   # exactly at the minimiser
    x0[1] = 0.0
    keep = ~np.isnan(orc.objective(ROSEN, x0, SEQ))      # the constructor asserts !isnan(f0)  (:773)
    x0 = np.ascontiguousarray(x0[keep])
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, L0, batched=True)
    ref = orc.BFGS(ROSEN, x0, L0, order=orc.SEQ, nthreads=8)
    _compare_state(opt, ref, True, "ctor")
    for it in range(10):
        dz.step_(opt); ref.step(1)
        _compare_state(opt, ref, True, f"n={n} L0={L0} iter {it}")
    opt.step(30); ref.step(30)
    _compare_state(opt, ref, True, "fused")


@pytest.mark.parametrize("n", [2050, 70, 260, 700])     # cluster / one warp per problem (1, 8 virtual warps) / cluster
@pytest.mark.parametrize("L0", [1.0, 1e-12, 1e6, 1e-300])
def test_large_awkward_inputs(gpu, orc, L0, n):
    dz = gpu
    EF = dz.ExampleFunctions
    u = orc.pcg_fill(n, 99)
    pick = (orc.pcg_fill(n, 98) * 40).astype(int)
    special = np.array([0.0, -0.0, 1.0, -1.0, 1e-8, -1e-8, 1e3, -1e3, 1e-160, 1e-300])
    x0 = np.where(pick < 10, special[np.minimum(pick, 9)], 4.0 * u - 2.0)
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, L0)
    ref = orc.BFGS(ROSEN, x0[None, :], L0, order=orc.TREE)
    _compare_state(opt, ref, False, "ctor")
    for it in range(10):
        dz.step_(opt); ref.step(1)
        _compare_state(opt, ref, False, f"L0={L0} iter {it}")


@pytest.mark.parametrize("n,batch", [(2, 1000), (16, 5000), (6, 777)])
def test_zero_copy_field_mirrors_equal_the_device_fields(gpu, orc, n, batch):
    """dzo_bfgs_mirror_fields: the step kernels store current_objective_value / has_terminated into page-locked host
    buffers while they run; the mirrored reads must equal ordinary copies of the device fields after every step!."""
    import ctypes as C
    dz = gpu
    EF = dz.ExampleFunctions
    x0 = (4.0 * orc.pcg_fill(n * batch, 31) - 2.0).reshape(batch, n)
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
    opt.reuse_host_buffers(True)
    ref = orc.BFGS(orc.OBJ_ROSENBROCK, x0, 1.0, order=orc.SEQ)
    f_copy, t_copy = np.empty(batch), np.empty(batch, dtype=np.uint8)
    for it in range(60):
        k = 1 if it < 40 else 7
        opt.step(k); ref.step(k)
        f_mirror = opt.current_objective_value           # only synchronises
        t_mirror = opt.has_converged
        assert dz.lib().dzo_bfgs_get_objective(opt._h, f_copy.ctypes.data_as(dz._capi.c_double_p)) == 0
        assert dz.lib().dzo_bfgs_get_terminated(opt._h, t_copy.ctypes.data_as(dz._capi.c_u8_p)) == 0
        assert_bitwise(f_mirror, f_copy, f"iter {it}: mirrored objective")
        assert np.array_equal(t_mirror.view(np.uint8), t_copy), f"iter {it}: mirrored has_terminated"
        assert_bitwise(f_mirror, ref.objective, f"iter {it}: objective vs oracle")
        assert np.array_equal(t_mirror, ref.terminated)
    assert t_copy.sum() > 0 or n == 16          # some problems have terminated (n = 2, 6 converge within the run)
    # pageable memory is refused, (NULL, NULL) clears the mirrors
    pageable = np.empty(batch)
    assert dz.lib().dzo_bfgs_mirror_fields(opt._h, pageable.ctypes.data_as(dz._capi.c_double_p), None) == -1
    assert dz.lib().dzo_bfgs_mirror_fields(opt._h, None, None) == 0
    opt._mirrored = False
    opt.step(1); ref.step(1)
    assert_bitwise(opt.current_objective_value, ref.objective, "after clearing the mirrors")


def test_field_mirrors_are_refused_where_no_kernel_maintains_them(gpu, orc):
    """Batched Riesz BFGS runs on the generic one-thread-per-problem kernel, which does not write mirrors: the C entry
    point refuses, the Python twin falls back to ordinary copies into its cached buffers, and the reads stay correct."""
    dz = gpu
    EF = dz.ExampleFunctions
    pts = (2.0 * orc.pcg_fill(40 * 6, 33) - 1.0).reshape(40, 2, 3)
    pts /= np.sqrt((pts * pts).sum(axis=2, keepdims=True))
    opt = dz.BFGSOptimizer(EF.riesz_energy, EF.riesz_gradient_, dz.SPHERE_CONSTRAINT, pts, 1e-2, batched=True)
    opt.reuse_host_buffers(True)
    assert opt._mirrored is False
    ref = orc.BFGS(orc.OBJ_RIESZ, pts.reshape(40, 6), 1e-2, order=orc.SEQ, constraint=orc.CONSTRAINT_SPHERE, dim=3)
    for it in range(5):
        opt.step(1); ref.step(1)
        assert_bitwise(opt.current_objective_value, ref.objective, f"iter {it}: objective")
        assert np.array_equal(opt.has_converged, ref.terminated)


# ----------------------------------------------------------------------------- round 2: implicit identities, step kinds, secant condition
def _oracle_kinds(ref, steps):
    """per-step kind counts replayed on the oracle: BFGS-type steps split by whether H was the identity going in
    (previous step gradient-descent or none: :781-783, :981), GD-type, terminating, idle."""
    out = []
    prev_type, prev_done = ref.step_type.copy(), ref.terminated.copy()
    for _ in range(steps):
        it0 = ref.iteration_count.copy()
        ref.step(1)
        moved = ref.iteration_count > it0
        ty, done = ref.step_type, ref.terminated
        ident_in = (prev_type != 2)
        out.append({"bfgs_read_h": int((moved & (ty == 2) & ~ident_in).sum()),
                    "bfgs_identity_h": int((moved & (ty == 2) & ident_in).sum()),
                    "gradient_descent": int((moved & (ty == 1)).sum()),
                    "terminate": int((done & ~prev_done).sum()),
                    "idle": int(prev_done.sum())})
        prev_type, prev_done = ty.copy(), done.copy()
    return out


@pytest.mark.parametrize("n,batch", [(16, 4000), (2, 3000), (8, 1000)])
def test_step_kind_counters_and_implicit_identity(gpu, orc, n, batch):
    """dzo_bfgs_get_step_kind_counts (what bench.py's roofline is computed from) equals the kinds replayed on the
    oracle; the inverse Hessian read through the C ABI is the oracle's even while the device keeps H = I implicit."""
    dz = gpu
    EF = dz.ExampleFunctions
    x0 = _x0(orc, n * batch, 515).reshape(batch, n)
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
    ref = orc.BFGS(ROSEN, x0, 1.0, order=orc.SEQ, nthreads=8)
    assert_bitwise(opt.inverse_hessian(3), np.eye(n), "H after the constructor (implicit identity)")
    total = {k: 0 for k in ("bfgs_read_h", "bfgs_identity_h", "gradient_descent", "terminate", "idle")}
    steps = 45
    want = _oracle_kinds(ref, steps)
    for it in range(steps):
        opt.step(1)
        got = opt.step_kind_counts()
        for k in total:
            total[k] += want[it][k]
        got_idle = got["idle"] + got["idle_warp"]
        assert {k: got[k] for k in total if k != "idle"} == {k: total[k] for k in total if k != "idle"}, f"iter {it}"
        assert got_idle == total["idle"], f"iter {it}: idle"
    _compare_state(opt, ref, True, "after the counted steps")
    for p in range(0, batch, max(1, batch // 40)):
        assert_bitwise(opt.inverse_hessian(p), ref.inverse_hessian(p), f"H[{p}]")
    assert total["gradient_descent"] > 0 and total["bfgs_identity_h"] > 0 and total["bfgs_read_h"] > 0
    assert sum(total.values()) == steps * batch
    # k fused steps in one launch are counted the same way
    opt.step_kind_counts(reset=True)
    more = _oracle_kinds(ref, 30)
    opt.step(30)
    got = opt.step_kind_counts()
    for k in ("bfgs_read_h", "bfgs_identity_h", "gradient_descent", "terminate"):
        assert got[k] == sum(m[k] for m in more), k
    assert got["idle"] + got["idle_warp"] == sum(m["idle"] for m in more)
    _compare_state(opt, ref, True, "after the fused counted steps")


def _secant_residual(H, dg, dx):
    """|H dg - dx|_inf / |dx|_inf  -- the secant condition the rank-2 update (:864-889) must restore"""
    return float(np.abs(H @ dg - dx).max() / np.abs(dx).max())


@pytest.mark.parametrize("n,batched", [(16, True), (2048, False)])
def test_secant_condition_after_every_bfgs_step(gpu, orc, n, batched):
    """An independent check of update_inverse_hessian! (:864-889) that neither restatement produced: after every
    BFGS-type step the new inverse Hessian maps delta_gradient to delta_point (H dg = dx, to rounding), and it stays
    bitwise symmetric.  Checked on the CUDA result itself (the oracle twin is tests/test_oracle.py)."""
    dz = gpu
    EF = dz.ExampleFunctions
    batch = 64 if batched else 1
    x0 = _x0(orc, n * batch, 612).reshape(batch, n)
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0 if batched else x0[0], 1.0, batched=batched)
    seen = 0
    for it in range(25 if batched else 12):
        opt.step(1)
        ty = np.atleast_1d(opt.last_step_type)
        moved = np.atleast_1d(opt.iteration_count) == it + 1
        dx = np.atleast_2d(opt.delta_point); dg = np.atleast_2d(opt.delta_gradient)
        for p in range(batch):
            if not (moved[p] and ty[p] == dz.StepType.BFGSStep):
                continue
            H = opt.inverse_hessian(p)
            assert_bitwise(H, H.T, f"iter {it} problem {p}: H symmetric")
            assert _secant_residual(H, dg[p], dx[p]) < 1e-10, f"iter {it} problem {p}: secant condition"
            seen += 1
    assert seen > (100 if batched else 5)


@pytest.mark.parametrize("n,batch", [(34, 50), (64, 40), (128, 33), (256, 12), (512, 6), (1100, 5)])
def test_batched_medium_n_trace(gpu, orc, n, batch):
    """README.md:12 "run multiple optimizers in parallel" for 32 < n: one handle, `batch` independent problems, every
    kernel instance (a thread-block cluster for the O(n) stage, blockIdx.z of the n^2 sweeps) serves one problem.  Same
    TREE summation order as a single large-n problem; every field of every problem equals the oracle's after every step!."""
    dz = gpu
    EF = dz.ExampleFunctions
    x0 = _x0(orc, n * batch, 4400 + n).reshape(batch, n)
    x0[1] = 1.0                                                  # one problem starts at the minimiser and terminates at once
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
    ref = orc.BFGS(ROSEN, x0, 1.0, order=orc.TREE, nthreads=8)
    assert opt.summation_order == TREE
    _compare_state(opt, ref, True, "ctor")
    for it in range(8):
        dz.step_(opt); ref.step(1)
        _compare_state(opt, ref, True, f"n={n} iter {it}")
    opt.step(5); ref.step(5)
    _compare_state(opt, ref, True, "after 5 more steps in one call")
    assert opt.count_active() == ref.count_active() < batch
    for p in (0, 1, batch - 1):
        assert_bitwise(opt.inverse_hessian(p), ref.inverse_hessian(p), f"H[{p}]")
    # a single-problem handle on problem 3 walks the same trajectory (the batch dimension changes no bit)
    one = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0[3], 1.0)
    one.step(13)
    assert_bitwise(one.current_point, opt.current_point[3], "batched vs single handle")
    # resume (set_state) of the whole batch
    H = np.stack([opt.inverse_hessian(p) for p in range(batch)])
    b = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
    b.set_state(opt.current_point, H, opt.delta_point, opt.delta_gradient, opt.last_step_length, opt.last_step_type,
                opt.iteration_count)
    # the oracle's own state-rebuilding constructor (:819-862): it recomputes d = H*g (:834-836), so a problem whose
    # d = copy(g) held a -0.0 (problem 1, terminated at the minimiser) legitimately resumes with +0.0
    Hr = np.stack([ref.inverse_hessian(p) for p in range(batch)])
    rb = orc.BFGS(ROSEN, x0, 1.0, order=orc.TREE, nthreads=8)
    rb.set_state(ref.point, Hr, ref.delta_point, ref.delta_gradient, ref.step_length, ref.step_type, ref.iteration_count)
    _compare_state(b, rb, True, "resumed batch vs the oracle's set_state")
    b.step(2); opt.step(2); ref.step(2); rb.step(2)
    _compare_state(b, rb, True, "resumed batch")
    _compare_state(opt, ref, True, "original batch")
    assert_bitwise(b.current_point, opt.current_point, "resumed trajectory")


@pytest.mark.parametrize("n,batch", [(64, 9), (200, 5), (512, 3)])
def test_warp_and_cluster_search_agree(gpu, orc, n, batch):
    """32 < n <= 512: the O(n) stage of step! runs on one warp per problem (warp_search.cuh); the 8-CTA cluster kernel
    (tuning knob warp_search = 0) must produce the same bits -- both are the oracle's DZO_ORDER_TREE."""
    dz = gpu
    EF = dz.ExampleFunctions
    x0 = _x0(orc, n * batch, 5100 + n).reshape(batch, n)
    x0[0, ::2] = -0.0                                            # signed zeros through the tree (+0.0 of the empty subtrees)
    ref = orc.BFGS(ROSEN, x0, 1.0, order=orc.TREE, nthreads=8)
    ref.step(9)
    try:
        for knob in (1, 0):
            dz.set_tuning("warp_search", knob)
            opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
            for _ in range(9):
                dz.step_(opt)
            _compare_state(opt, ref, True, f"warp_search={knob}")
    finally:
        dz.set_tuning("warp_search", 1)
