"""update_gemv_kernel / gemv_kernel at n = 16384 with the default tuning (fractions of the HBM copy peak), best of 5."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dzopt_b200 as dz
n = 16384
peak, _ = bench.load_peaks()
ms = C.c_float()
row = {"lib": os.path.basename(os.environ.get("DZOPT_B200_LIB", "default"))}
for which, name, nbytes in ((1, "gemv", 8), (2, "update_gemv", 16)):
    best = None
    for _ in range(5):
        assert dz.lib().dzo_bench_kernel(which, n, 10, 0, C.byref(ms), 0) == 0
        best = ms.value if best is None else min(best, ms.value)
    row[name] = round(nbytes * n * n / (best * 1e-3) / 1e9 / peak, 4)
print(json.dumps(row), flush=True)
