"""Pairwise radial N-body kernels of the live package (src/ExampleFunctions.jl:16-72, :117-468):
CPU tests of the oracle (C vs the independent Python restatement, physics checks) and GPU parity tests."""
import numpy as np
import pytest

from conftest import assert_bitwise

SEQ, TREE = 0, 1


def cloud(orc, n, seed, spread=None):
    """n particles, PCG-uniform in a box whose volume grows with n (so distances stay O(1))."""
    L = spread if spread is not None else 1.2 * n ** (1.0 / 3.0)
    p = orc.pcg_fill(3 * n, seed).reshape(3, n) * L
    return p[0].copy(), p[1].copy(), p[2].copy()


def lattice(orc, n, seed):
    """jittered cubic lattice (spacing 1.1): no close pairs, so energies stay O(n) and finite differences work"""
    m = int(np.ceil(n ** (1.0 / 3.0)))
    idx = np.arange(m ** 3)[:n]
    base = np.stack([idx % m, (idx // m) % m, idx // (m * m)]).astype(np.float64) * 1.1
    p = base + 0.1 * (orc.pcg_fill(3 * n, seed).reshape(3, n) - 0.5)
    return p[0].copy(), p[1].copy(), p[2].copy()


# ----------------------------------------------------------------------------- CPU: oracle
@pytest.mark.parametrize("n,tree", [(1, False), (2, False), (37, False), (130, True), (300, True)])
def test_c_oracle_equals_python_restatement(orc, n, tree):
    import dzo_oracle_py as P
    x, y, z = cloud(orc, n, 5)
    u, v, w = (a - 0.5 for a in cloud(orc, n, 6, 1.0))
    order = orc.TREE if tree else orc.SEQ
    pe, e = orc.pairwise_energy(x, y, z, order)
    assert_bitwise(pe, np.array(P.pairwise_point_energies(list(x), list(y), list(z), tree)), "point energies")
    assert e == P.pairwise_energy(list(x), list(y), list(z), tree)
    for a, b in zip(orc.pairwise_gradient(x, y, z, order), P.pairwise_gradient(list(x), list(y), list(z), tree)):
        assert_bitwise(a, np.array(b), "gradient")
    for a, b in zip(orc.pairwise_hvp(x, y, z, u, v, w, order),
                    P.pairwise_hvp(list(x), list(y), list(z), list(u), list(v), list(w), tree)):
        assert_bitwise(a, np.array(b), "hvp")


def test_lennard_jones_physics(orc):
    """dimer at the LJ minimum r = 2^(1/6): energy -1, zero force; the CPU twin (pairwise_radial_energy,
    src/ExampleFunctions.jl:83-114: sum over i<j) equals the kernel's half-double-counted sum."""
    r = 2.0 ** (1.0 / 6.0)
    x, y, z = np.array([0.0, r]), np.zeros(2), np.zeros(2)
    pe, e = orc.pairwise_energy(x, y, z)
    assert abs(e + 1.0) < 1e-15 and abs(pe[0] + 0.5) < 1e-15
    g = orc.pairwise_gradient(x, y, z)
    assert max(np.abs(a).max() for a in g) < 1e-13
    x, y, z = cloud(orc, 60, 7)
    _, e = orc.pairwise_energy(x, y, z)
    ref = 0.0
    for i in range(60):
        d2 = (x[i] - x[i + 1:]) ** 2 + (y[i] - y[i + 1:]) ** 2 + (z[i] - z[i + 1:]) ** 2
        ref += (4.0 * (d2 ** -6 - d2 ** -3)).sum()
    assert abs(e - ref) <= 1e-12 * abs(ref)


def test_gradient_and_hvp_are_derivatives(orc):
    """finite differences (legacy/ExampleFunctions.jl:290-303 style): g = dE/dp, Hv = d/dt g(p + t v)."""
    n = 24
    x, y, z = lattice(orc, n, 8)
    E = lambda a, b, c: orc.pairwise_energy(a, b, c)[1]
    g = orc.pairwise_gradient(x, y, z)
    h = 1e-6
    for k, arr in enumerate((x, y, z)):
        for i in (0, 5, n - 1):
            p, m = [x.copy(), y.copy(), z.copy()], [x.copy(), y.copy(), z.copy()]
            p[k][i] += h; m[k][i] -= h
            fd = (E(*p) - E(*m)) / (2 * h)
            assert abs(fd - g[k][i]) <= 1e-5 * max(1.0, abs(fd))
    u, v, w = (a - 0.5 for a in cloud(orc, n, 9, 1.0))
    hv = orc.pairwise_hvp(x, y, z, u, v, w)
    gp = orc.pairwise_gradient(x + h * u, y + h * v, z + h * w)
    gm = orc.pairwise_gradient(x - h * u, y - h * v, z - h * w)
    for k in range(3):
        fd = (gp[k] - gm[k]) / (2 * h)
        assert np.abs(fd - hv[k]).max() <= 1e-4 * max(1.0, np.abs(fd).max())


def test_tree_and_sequential_agree_to_rounding(orc):
    x, y, z = cloud(orc, 700, 10)
    a, b = orc.pairwise_gradient(x, y, z, orc.SEQ), orc.pairwise_gradient(x, y, z, orc.TREE)
    for p, q in zip(a, b):
        assert np.abs(p - q).max() <= 1e-12 * np.abs(p).max()


# ----------------------------------------------------------------------------- GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 31, 256, 257, 1000, 4096])
@pytest.mark.parametrize("order", [SEQ, TREE])
def test_gpu_pairwise_matches_oracle_bitwise(gpu, orc, n, order):
    dz = gpu
    EF = dz.ExampleFunctions
    x, y, z = cloud(orc, n, 11)
    u, v, w = (a - 0.5 for a in cloud(orc, n, 12, 1.0))
    e = dz.accelerated_pairwise_radial_energy(EF.lj_energy, x, y, z, order=order)
    assert e == orc.pairwise_energy(x, y, z, order)[1]
    gx, gy, gz = np.empty(n), np.empty(n), np.empty(n)
    assert dz.accelerated_pairwise_radial_gradient_(gx, gy, gz, EF.lj_first_derivative, x, y, z, order=order) is None
    for a, b, nm in zip((gx, gy, gz), orc.pairwise_gradient(x, y, z, order), "xyz"):
        assert_bitwise(a, b, f"g{nm}")
    px, py, pz = np.empty(n), np.empty(n), np.empty(n)
    dz.accelerated_pairwise_radial_hvp_(px, py, pz, EF.lj_first_derivative, EF.lj_second_derivative, x, y, z, u, v, w, order=order)
    for a, b, nm in zip((px, py, pz), orc.pairwise_hvp(x, y, z, u, v, w, order), "xyz"):
        assert_bitwise(a, b, f"p{nm}")


@pytest.mark.gpu
def test_gpu_pairwise_device_arrays_async(gpu, orc):
    """the reference's calling convention: device arrays in, device arrays out, no synchronisation"""
    import torch
    dz = gpu
    EF = dz.ExampleFunctions
    n = 3000
    x, y, z = cloud(orc, n, 13)
    tx, ty, tz = (torch.from_numpy(a).cuda() for a in (x, y, z))
    g = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)]
    dz.accelerated_pairwise_radial_gradient_(*g, EF.lj_first_derivative, tx, ty, tz, order=TREE)
    e = dz.accelerated_pairwise_radial_energy(EF.lj_energy, tx, ty, tz, order=TREE)
    torch.cuda.synchronize()
    for a, b in zip(g, orc.pairwise_gradient(x, y, z, orc.TREE)):
        assert_bitwise(a.cpu().numpy(), b, "device gradient")
    assert e == orc.pairwise_energy(x, y, z, orc.TREE)[1]
    with pytest.raises(TypeError):
        dz.accelerated_pairwise_radial_energy(EF.lj_first_derivative, tx, ty, tz)
    with pytest.raises(AssertionError):
        dz.accelerated_pairwise_radial_energy(EF.lj_energy, tx, ty[:-1], tz)
