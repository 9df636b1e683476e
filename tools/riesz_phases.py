"""Phase breakdown of one GD step! on Riesz N=4096 from the kernel's own leader-thread log (riesz_profile knob)."""
import os, sys, ctypes as C
from collections import defaultdict
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dzopt_b200 as dz
import oracle as orc
EF = dz.ExampleFunctions
if os.environ.get('DZO_RIESZ_BAR'):
    dz.set_tuning('riesz_bar', int(os.environ['DZO_RIESZ_BAR']))
if os.environ.get('DZO_RIESZ_THREADS'):
    dz.set_tuning('riesz_threads', int(os.environ['DZO_RIESZ_THREADS']))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
p = 2.0 * orc.pcg_fill(3 * N, 3).reshape(N, 3) - 1.0
p = p / np.sqrt((p * p).sum(axis=1, keepdims=True))
if os.environ.get("DZO_RIESZ_PAIR"):
    dz.set_tuning("riesz_pair", int(os.environ["DZO_RIESZ_PAIR"]))
o = dz.GradientDescentOptimizer(dz.SPHERE_CONSTRAINT, EF.riesz_energy, EF.riesz_gradient_, dz.QuadraticLineSearch(0), p, 1e-3)
o.step(3)
dz.set_tuning("riesz_profile", 1)
o.step(k)
ev = np.zeros(2 * 8192, dtype=np.uint64)
cnt = C.c_int64()
assert dz.lib().dzo_gd_get_phase_log(o._h, ev.ctypes.data_as(C.POINTER(C.c_uint64)), 8192, C.byref(cnt)) == 0
ev = ev[:2 * cnt.value].reshape(-1, 2)
if os.environ.get("DZO_RIESZ_PAIR"):
    dz.set_tuning("riesz_pair", int(os.environ["DZO_RIESZ_PAIR"]))
names = {(1, 12): "line search bookkeeping before a paired energy", (5, 12): "bookkeeping between energies (next: paired)",
         (12, 3): "paired energy: leader's own items", (1, 2): "line search bookkeeping before an energy", (5, 2): "bookkeeping between energies", (2, 3): "energy: leader's own items",
         (3, 4): "energy: wait + barrier", (4, 5): "energy: tree over the rows (per CTA)", (5, 6): "line search tail",
         (6, 7): "point update + barrier", (7, 8): "gradient: leader's own items", (8, 10): "gradient: wait + barrier",
         (10, 11): "dots + direction", (10, 13): "both norms (per CTA)", (13, 11): "direction + control block", (11, 1): "hand-off to the next step"}
tot, num = defaultdict(float), defaultdict(int)
for (a, ta), (b, tb) in zip(ev[:-1], ev[1:]):
    key = (int(a), int(b))
    tot[key] += float(tb - ta) * 1e-3
    num[key] += 1
span = float(ev[-1, 1] - ev[0, 1]) * 1e-3
print(f"{cnt.value} events over {k} steps, {span / k:.1f} us per step, {sum(1 for e in ev if e[0] == 2) / k:.2f} single + {sum(1 for e in ev if e[0] == 12) / k:.2f} paired energy phases per step")
for key in sorted(tot, key=lambda q: -tot[q]):
    print(f"  {names.get(key, str(key)):45s} {tot[key] / k:8.2f} us/step  ({num[key] / k:.2f} x {tot[key] / num[key]:.2f} us)")
