"""The three n^2 kernels of a large-n BFGS step! timed alone (dzo_bench_kernel) over threads-per-CTA x columns-in-flight."""
import ctypes as C
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dzopt_b200 as dz

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
peak, _ = bench.load_peaks()
ms = C.c_float()
for threads in (64, 128, 256):
    for unroll in (8, 16, 24, 32):
        dz.set_tuning("sweep_threads", threads)
        dz.set_tuning("sweep_unroll", unroll)
        row = {"n": n, "sweep_threads": threads, "sweep_unroll": unroll}
        for which, name, nbytes in ((1, "gemv", 8), (2, "update_gemv", 16), (3, "identity", 8)):
            best = None
            for _ in range(3):
                rc = dz.lib().dzo_bench_kernel(which, n, 10, 0, C.byref(ms), 0)
                assert rc == 0, dz.lib().dzo_last_error()
                best = ms.value if best is None else min(best, ms.value)
            row[name] = round(nbytes * n * n / (best * 1e-3) / 1e9 / peak, 4)
        print(json.dumps(row), flush=True)
