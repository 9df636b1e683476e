"""legacy/PCG.jl:7-22 -- the reference's synthetic-input generator (``PCG.random_fill!``), host side, numpy.

The 64-bit LCG is sequential; it is evaluated in columns: the first ``B`` states one by one, then every later row of
``B`` states from the previous row through the B-step affine jump (A^B, C_B), vectorised over the row.  Bit-identical
to ``dzo_cpu_pcg_fill`` of the oracle (tests/test_abi_and_host.py), but owned by the product so that ``bench.py`` and
the tools generate their inputs without touching ``oracle/``."""
import numpy as np

_A = 0x5851F42D4C957F2D
_C = 0x14057B7EF767814F
_M = (1 << 64) - 1


def pcg_fill(count: int, seed: int) -> np.ndarray:
    """``PCG.random_fill!(zeros(count), seed)``: ``count`` doubles, uniform in [0, 1) with 32 random bits."""
    count = int(count)
    if count <= 0:
        return np.empty(0)
    B = 4096 if count > 4096 else count
    rows = (count + B - 1) // B
    first = np.empty(B, dtype=np.uint64)
    s = (_A * ((_C + int(seed)) & _M) + _C) & _M               # :16  state = advance(INCREMENT + seed)
    aj, cj = 1, 0                                             # affine map of j steps: s -> aj*s + cj
    for i in range(B):
        first[i] = s
        s = (_A * s + _C) & _M                                # :7-8
        aj, cj = (_A * aj) & _M, (_A * cj + _C) & _M
    states = np.empty((rows, B), dtype=np.uint64)
    states[0] = first
    a_b, c_b = np.uint64(aj), np.uint64(cj)
    with np.errstate(over="ignore"):
        for r in range(1, rows):
            states[r] = states[r - 1] * a_b + c_b             # wraps mod 2^64
    st = states.reshape(-1)[:count]
    v = (((st >> np.uint64(18)) ^ st) >> np.uint64(27)).astype(np.uint32)       # :11-12
    k = (st >> np.uint64(59)).astype(np.uint32)
    out = (v >> k) | (v << ((np.uint32(32) - k) & np.uint32(31)))               # bitrotate right by the top 5 bits
    return out.astype(np.float64) * 2.3283064365386962890625e-10               # :18  2^-32
