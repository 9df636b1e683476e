"""Property tests (hypothesis): the two independent restatements of the reference path -- oracle/dzo_oracle.c
and oracle/dzo_oracle_py.py -- must agree BIT FOR BIT on arbitrary small inputs, including the awkward ones
(zeros, huge / tiny magnitudes, tiny or enormous initial step lengths, points already at the minimum)."""
import math

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import assert_bitwise

ROSEN, RIESZ = 1, 2

coord = st.one_of(
    st.floats(min_value=-3.0, max_value=3.0, allow_nan=False),
    st.sampled_from([0.0, -0.0, 1.0, -1.0, 1e-8, -1e-8, 1e6, -1e6, 1e-160, 1e150]),
)
step_len = st.sampled_from([1.0, 0.5, 1e-3, 1e-12, 37.0, 1e6, 1e-300])
COMMON = dict(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@settings(**COMMON)
@given(data=st.data())
def test_bfgs_restatements_agree(orc, data):
    import dzo_oracle_py as P
    n = 2 * data.draw(st.integers(min_value=1, max_value=5))
    tree = data.draw(st.booleans())
    x0 = [data.draw(coord) for _ in range(n)]
    L0 = data.draw(step_len)
    f0 = P.Rosenbrock(tree).f(list(x0))
    if math.isnan(f0):
        with pytest.raises(orc.OracleError):
            orc.BFGS(ROSEN, np.array([x0]), L0, order=orc.TREE if tree else orc.SEQ)
        return
    py = P.BFGSOptimizer(P.Rosenbrock(tree), list(x0), L0, tree)
    c = orc.BFGS(ROSEN, np.array([x0]), L0, order=orc.TREE if tree else orc.SEQ)
    for it in range(12):
        py.step(); c.step(1)
        assert_bitwise(c.point[0], np.array(py.current_point), f"iter {it} point")
        assert_bitwise(c.direction[0], np.array(py.next_step_direction), f"iter {it} direction")
        assert_bitwise(c.inverse_hessian(0), np.array(py.H), f"iter {it} H")
        assert_bitwise(c.objective, np.array([py.current_objective_value]), "objective")
        assert_bitwise(c.step_length, np.array([py.last_step_length]), "step length")
        assert int(c.step_type[0]) == py.last_step_type and bool(c.terminated[0]) == py.has_terminated
        if py.has_terminated:
            break


@settings(**COMMON)
@given(data=st.data())
def test_gd_riesz_restatements_agree(orc, data):
    import dzo_oracle_py as P
    dim = data.draw(st.integers(min_value=1, max_value=3))
    N = data.draw(st.integers(min_value=2, max_value=6))
    sphere = data.draw(st.booleans())
    tree = data.draw(st.booleans())
    pts = [data.draw(st.floats(min_value=-2.0, max_value=2.0, allow_nan=False)) for _ in range(dim * N)]
    # the sphere constraint divides by the norm of every point: keep them away from the origin
    for j in range(N):
        if sum(c * c for c in pts[j * dim:(j + 1) * dim]) < 1e-6:
            pts[j * dim] = 1.0 + j
    L0 = data.draw(st.sampled_from([1e-2, 1.0, 1e-9]))
    mi = data.draw(st.integers(min_value=0, max_value=3))
    py = P.GradientDescentOptimizer(P.Riesz(dim, sphere, tree), list(pts), L0, mi, tree)
    c = orc.GD(RIESZ, np.array([pts]), L0, order=orc.TREE if tree else orc.SEQ, constraint=int(sphere), dim=dim, max_increases=mi)
    for it in range(6):
        py.step(); c.step(1)
        assert_bitwise(c.point[0], np.array(py.current_point), f"iter {it} point")
        assert_bitwise(c.direction[0], np.array(py.next_step_direction), f"iter {it} direction")
        assert_bitwise(c.objective, np.array([py.current_objective_value]), "objective")
        assert bool(c.terminated[0]) == py.has_terminated
        if py.has_terminated:
            break


@settings(**COMMON)
@given(data=st.data())
def test_line_search_restatements_agree(orc, data):
    import dzo_oracle_py as P
    n = 2 * data.draw(st.integers(min_value=1, max_value=4))
    x = [data.draw(coord) for _ in range(n)]
    d = [data.draw(coord) for _ in range(n)]
    t1 = data.draw(st.sampled_from([1.0, 1e-6, 1e6, 0.0, float("inf"), 1e-320, 3.0]))
    fn = P.Rosenbrock(False)
    f0 = fn.f(list(x))
    ref = P.quadratic_line_search(P.Ray(fn, list(x), list(d), -1.0), f0, t1, 0)
    got = orc.line_search(ROSEN, np.array(x), np.array(d), f0, t1, orc.SEQ)
    assert (np.array(got).view(np.uint64) == np.array(ref, dtype=np.float64).view(np.uint64)).all() or (
        all(math.isnan(a) and math.isnan(b) or a == b for a, b in zip(got, ref)))


@settings(**{**COMMON, "max_examples": 12})
@given(data=st.data())
def test_legacy_lbfgs_restatements_agree(orc, data):
    """legacy LBFGSOptimizer (legacy/DZOptimization.jl:458-695) with arbitrary decorators, history lengths and starts"""
    import dzo_oracle_py as P
    n = 2 * data.draw(st.integers(min_value=1, max_value=5))
    tree = data.draw(st.booleans())
    x0 = [data.draw(coord) for _ in range(n)]
    L0 = data.draw(step_len)
    m = data.draw(st.integers(min_value=1, max_value=4))
    mi = data.draw(st.integers(min_value=0, max_value=3))
    lam = data.draw(st.sampled_from([None, 0.0, 0.25, 1e3]))
    box = data.draw(st.sampled_from([None, (-0.5, 0.8), (0.0, 0.0), (-1e9, 1e9), (1.0, 1.0)]))
    fn = P.Rosenbrock(tree)
    if lam is not None:
        fn = P.L2Regularized(fn, lam)
    if box is not None:
        fn = P.UniformBox(fn, *box)
    py = P.LegacyLBFGSOptimizer(fn, list(x0), L0, m, mi, tree)
    c = orc.LegacyLBFGS(ROSEN, np.array(x0), L0, m, mi, lam, box, orc.TREE if tree else orc.SEQ)

    def same(a, b):     # NaN payloads are not part of the contract
        a, b = np.array(a, dtype=np.float64), np.array(b, dtype=np.float64)
        assert ((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))).all()

    for it in range(6):
        py.step(); c.step(1)
        same(c.point, py.current_point)
        same(c.direction, py.next_step_direction)
        same(c.gradient, py.current_gradient)
        s = c.scalars
        same(s[:3], [py.current_objective_value, py.delta_objective_value, py.last_step_length])
        assert int(s[3]) == py.iteration_count and bool(s[4]) == py.has_terminated and int(s[5]) == py._history_count
        rho, alpha = c.history
        same(rho, py._rho); same(alpha, py._alpha)
        if py.has_terminated:
            break
