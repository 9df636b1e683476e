import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/dzo_oracle.c through oracle/oracle.py) -- the checker."""
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def dz():
    """The product package (host mirror of the Julia API over libdzopt_b200.so)."""
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build_cuda()
    import dzopt_b200
    return dzopt_b200


@pytest.fixture(scope="session")
def gpu(dz):
    """Fails loudly (never skips to a fallback) when the CUDA library cannot reach a device."""
    import ctypes as C
    out = C.c_double()
    v = np.ones(2)
    rc = dz.lib().dzo_dot(0, 2, v.ctypes.data_as(dz._capi.c_double_p), v.ctypes.data_as(dz._capi.c_double_p),
                          C.byref(out), 0)
    assert rc == 0, dz.lib().dzo_last_error().decode()
    assert out.value == 2.0
    return dz


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bitwise(a, b, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    same = bits(a) == bits(b)
    if not same.all():
        idx = np.argwhere(~same)[0]
        raise AssertionError(f"{what}: {int((~same).sum())} of {same.size} values differ bitwise; first at "
                             f"{tuple(idx)}: {a[tuple(idx)]!r} vs {b[tuple(idx)]!r}")
