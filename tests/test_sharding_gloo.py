"""CPU model of the row-sharded large-n BFGS step! with world_size = 2 over gloo.

Each rank owns a row slab of H and replicated vectors, runs the O(n) line search redundantly, computes
its rows of t = H*dg and d = H'*g and all-gathers them -- exactly the data flow of
dzo_bfgs_create_sharded / large_step_once in csrc/dzopt_bfgs.cu, with the oracle's kernels standing in
for the CUDA ones and gloo for NCCL.  The result must equal the UNSHARDED oracle bit for bit, i.e. the
sharding changes no rounding and no control flow."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROSEN = 1


def _worker(rank, world, port, n, steps, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import oracle as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = orc.lib()
    dp = lambda a: a.ctypes.data_as(orc._capi.c_double_p)
    T = orc.TREE
    r0, r1 = rank * n // world, (rank + 1) * n // world
    rows = r1 - r0

    def allgather(local):
        parts = [torch.empty(rows, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(local)))
        return torch.cat(parts).numpy()

    # constructor :762-810 (replicated vectors, H slab = rows r0..r1 of I, column-major rows x n)
    x = 4.0 * orc.pcg_fill(n, 2) - 2.0
    f = orc.objective(ROSEN, x, T)[0]
    g = orc.gradient(ROSEN, x, T)[0]
    d = g.copy()
    L = 1.0
    H = np.zeros((n, rows))                      # H[j, i] = slab element (local row i, column j)
    H[np.arange(r0, r1), np.arange(rows)] = 1.0
    ref = orc.BFGS(ROSEN, x[None, :], 1.0, order=T)
    types = []
    for it in range(steps):
        ref.step(1)
        # step! :891-994, every rank identically
        gn = np.sqrt(orc.dot(g, g, T)); dn = np.sqrt(orc.dot(d, d, T))
        tg, fg = orc.line_search(ROSEN, x, g, f, L / gn, T)
        tb, fb = orc.line_search(ROSEN, x, d, f, L / dn, T)
        if fb < f and not (fb > fg):
            kind, alpha, direction, f, L = 2, -tb, d, fb, tb * dn
        elif fg < f:
            kind, alpha, direction, f, L = 1, -tg, g, fg, tg * gn
        else:
            types.append(0)
            break
        types.append(kind)
        xn = x + alpha * direction
        gnew = orc.gradient(ROSEN, xn, T)[0]
        dg = (-g) + gnew
        x, g = xn, gnew
        if kind == 2:
            overlap = orc.dot(direction, dg, T)
            sd = direction * (1.0 / overlap)
            t_loc = np.empty(rows)
            assert lib.dzo_cpu_gemv_rows(T, n, r0, r1, dp(H), dp(dg), dp(t_loc)) == 0
            t = allgather(t_loc)                                   # ncclAllGather #1
            delta = alpha * overlap + orc.dot(dg, t, T)            # recomputed redundantly after the gather
            assert lib.dzo_cpu_update_rows(n, r0, r1, dp(H), C.c_double(delta), dp(sd), dp(t)) == 0
            d_loc = np.empty(rows)
            assert lib.dzo_cpu_gemv_rows(T, n, r0, r1, dp(H), dp(g), dp(d_loc)) == 0
            d = allgather(d_loc)                                   # ncclAllGather #2
        else:
            H[:] = 0.0
            H[np.arange(r0, r1), np.arange(rows)] = 1.0
            d = g.copy()
        same = (np.array_equal(x.view(np.uint64), ref.point[0].view(np.uint64))
                and np.array_equal(d.view(np.uint64), ref.direction[0].view(np.uint64))
                and f == ref.objective[0] and L == ref.step_length[0])
        Href = ref.inverse_hessian(0)[r0:r1]                        # (rows, n)
        same = same and np.array_equal(np.ascontiguousarray(H.T).view(np.uint64), np.ascontiguousarray(Href).view(np.uint64))
        if not same:
            out.put((rank, f"mismatch at iteration {it}"))
            dist.destroy_process_group()
            return
    out.put((rank, "ok " + str(types)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,steps", [(64, 12), (2050, 4)])
def test_row_sharded_model_equals_unsharded_oracle(orc, n, steps):
    world = 2
    if n % (2 * world):
        n += 2 * world - n % (2 * world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg.startswith("ok"), f"rank {rank}: {msg}"
    assert "2" in results[0][1]            # at least one BFGS-type step exercised the gathers
