"""Thin numpy wrappers over the kernel-level C-ABI entry points (dzo_dev_*), for the GPU tests."""
import ctypes as C

import numpy as np

import dzopt_b200 as dz

_dp = lambda a: a.ctypes.data_as(dz._capi.c_double_p)


def _check(rc):
    if rc != 0:
        raise dz.DZOptError(rc, dz.lib().dzo_last_error().decode())


def dot(v, w, order):
    v = np.ascontiguousarray(v, dtype=np.float64); w = np.ascontiguousarray(w, dtype=np.float64)
    out = C.c_double()
    _check(dz.lib().dzo_dot(order, v.size, _dp(v), _dp(w), C.byref(out), 0))
    return out.value


def gemv(H, v, order):
    n = H.shape[0]
    Hc = np.ascontiguousarray(H.T, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.empty(n)
    _check(dz.lib().dzo_gemv(order, n, _dp(Hc), _dp(v), _dp(out), 0))
    return out


def update_inverse_hessian(H, step_length, d, dg, next_g, order):
    n = H.shape[0]
    Hc = np.ascontiguousarray(H.T, dtype=np.float64)
    d = np.array(d, dtype=np.float64); dg = np.ascontiguousarray(dg, dtype=np.float64)
    scratch = np.empty(n)
    if next_g is None:
        _check(dz.lib().dzo_update_inverse_hessian(order, n, _dp(Hc), float(step_length), _dp(d), _dp(dg),
                                                   _dp(scratch), None, None, 0))
        nd = None
    else:
        ng = np.ascontiguousarray(next_g, dtype=np.float64); nd = np.empty(n)
        _check(dz.lib().dzo_update_inverse_hessian(order, n, _dp(Hc), float(step_length), _dp(d), _dp(dg),
                                                   _dp(scratch), _dp(ng), _dp(nd), 0))
    return Hc.T, d, scratch, nd


def identity(n):
    H = np.empty((n, n))
    _check(dz.lib().dzo_identity(n, _dp(H), 0))
    return H


def objective(obj, x, order, constraint=0, dim=0):
    a = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
    f = np.empty(a.shape[0])
    _check(dz.lib().dzo_objective(obj, constraint, dim, order, a.shape[1], a.shape[0], _dp(a), _dp(f), 0))
    return f


def gradient(obj, x, order, constraint=0, dim=0):
    a = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
    g = np.empty_like(a)
    _check(dz.lib().dzo_gradient(obj, constraint, dim, order, a.shape[1], a.shape[0], _dp(a), _dp(g), 0))
    return g


def line_search(obj, x, direction, f0, t1, order, constraint=0, dim=0):
    x = np.ascontiguousarray(x, dtype=np.float64); d = np.ascontiguousarray(direction, dtype=np.float64)
    tb, fb = C.c_double(), C.c_double()
    _check(dz.lib().dzo_line_search(obj, constraint, dim, order, x.size, _dp(x), _dp(d), float(f0), float(t1),
                                    C.byref(tb), C.byref(fb), 0))
    return tb.value, fb.value


def bench_kernel(which, n, reps, variant=0):
    ms = C.c_float()
    _check(dz.lib().dzo_bench_kernel(which, n, reps, variant, C.byref(ms), 0))
    return ms.value
