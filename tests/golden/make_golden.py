"""Generate tests/golden/*.json from the independent pure-Python restatement (oracle/dzo_oracle_py.py).

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors and cannot run here (no Julia, and the
path is commented-out code with undefined helpers), so these fixtures pin the *restated* algorithm:
they are produced by one restatement (Python floats) and checked bit-for-bit against the other
(oracle/dzo_oracle.c) on the CPU and against the CUDA library on the GPU.  Every float is stored
as float.hex() so the comparison is exact.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import dzo_oracle_py as P  # noqa: E402

hx = lambda v: [float(a).hex() for a in v]


def bfgs_trace(x0, step, tree, iters, keep_points=True):
    opt = P.BFGSOptimizer(P.Rosenbrock(tree), x0, step, tree)
    rows = []
    for _ in range(iters):
        opt.step()
        row = {"f": opt.current_objective_value.hex(), "L": float(opt.last_step_length).hex(),
               "type": opt.last_step_type, "iter": opt.iteration_count, "term": opt.has_terminated}
        if keep_points:
            row["x"] = hx(opt.current_point)
        rows.append(row)
        if opt.has_terminated:
            break
    return {"x0": hx(x0), "step": step, "tree": tree, "rows": rows, "final_x": hx(opt.current_point),
            "final_d": hx(opt.next_step_direction), "final_H_row0": hx(opt.H[0])}


def gd_trace(fn, x0, step, tree, iters, max_increases=0):
    opt = P.GradientDescentOptimizer(fn, x0, step, max_increases, tree)
    rows = []
    for _ in range(iters):
        opt.step()
        rows.append({"f": float(opt.current_objective_value).hex(), "L": float(opt.last_step_length).hex(),
                     "iter": opt.iteration_count, "term": opt.has_terminated})
        if opt.has_terminated:
            break
    return {"x0": hx(x0), "step": step, "tree": tree, "max_increases": max_increases, "rows": rows,
            "final_x": hx(opt.current_point), "final_d": hx(opt.next_step_direction)}


def sphere_points(N, dim, seed):
    u = P.pcg_fill(N * dim, seed)
    out = []
    for j in range(N):
        p = [2.0 * u[j * dim + k] - 1.0 for k in range(dim)]
        s = sum(c * c for c in p) ** 0.5
        out += [c / s for c in p]
    return out


def main():
    g = {}
    g["pcg"] = {str(seed): hx(P.pcg_fill(8, seed)) for seed in (0, 1, 2024, 2 ** 63 + 5)}
    # config 1: README Rosenbrock n=2 from "rand(2)" (PCG seeds), step 1.0, run to has_converged
    g["c1_rosenbrock_n2"] = [bfgs_trace(P.pcg_fill(2, s), 1.0, False, 500) for s in range(6)]
    # config 2 (one problem of the batch): n=16, x0 = 4u-2, seed 2024
    u = P.pcg_fill(48, 2024)
    g["c2_rosenbrock_n16"] = [bfgs_trace([4.0 * a - 2.0 for a in u[16 * p:16 * p + 16]], 1.0, False, 40) for p in range(3)]
    # large-n path, TREE order: n=64 (one GEMV chunk) and n=1100 (two chunks)
    u = P.pcg_fill(64, 1)
    g["tree_rosenbrock_n64"] = bfgs_trace([4.0 * a - 2.0 for a in u], 1.0, True, 25)
    u = P.pcg_fill(1100, 1)
    g["tree_rosenbrock_n1100"] = bfgs_trace([4.0 * a - 2.0 for a in u], 1.0, True, 4, keep_points=False)
    # kernel-level: tree dot, tree gemv on a non-symmetric matrix
    n = 1030
    a = P.pcg_fill(n * 3, 77)
    g["tree_dot_n1030"] = {"seed": 77, "value": P.dot(a[:n], a[n:2 * n], True).hex()}
    rows = 6
    m = P.pcg_fill(n * rows, 78)
    H = [[m[i * n + j] - 0.5 for j in range(n)] for i in range(rows)]
    v = a[2 * n:3 * n]
    out = []
    for i in range(rows):
        acc = None
        for c0 in range(0, n, P.GEMV_CHUNK):
            part = 0.0
            for j in range(c0, min(c0 + P.GEMV_CHUNK, n)):
                part += H[i][j] * v[j]
            acc = part if acc is None else acc + part
        out.append(acc)
    g["tree_gemv_rows_n1030"] = {"seed_matrix": 78, "seed_vector": 77, "rows": rows, "values": hx(out)}
    # GD: Riesz on the sphere (config 5 in miniature: 40 points and 300 points = 3 segments), Rosenbrock
    g["gd_riesz_sphere_N40"] = gd_trace(P.Riesz(3, True, True), sphere_points(40, 3, 3), 1e-3, True, 12)
    g["gd_riesz_sphere_N300"] = gd_trace(P.Riesz(3, True, True), sphere_points(300, 3, 3), 1e-3, True, 2)
    g["gd_riesz_free_N20_seq"] = gd_trace(P.Riesz(2, False, False), sphere_points(20, 2, 5), 1e-2, False, 10, 3)
    u = P.pcg_fill(64, 6)
    g["gd_rosenbrock_n64"] = gd_trace(P.Rosenbrock(True), [4.0 * a - 2.0 for a in u], 1e-2, True, 20)
    # objectives
    pts = sphere_points(300, 3, 61)
    grad = [0.0] * len(pts)
    P.Riesz(3, True, True).g(grad, pts)
    g["riesz_N300"] = {"seed": 61, "energy_tree": P.Riesz(3, True, True).f(pts).hex(),
                       "energy_seq": P.Riesz(3, True, False).f(pts).hex(), "gradient_tree_first6": hx(grad[:6])}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=0, sort_keys=True)
    print("wrote golden.json:", {k: (len(v) if isinstance(v, list) else "ok") for k, v in g.items()})


def legacy_trace(x0, step, m, mi, lam, box, tree, iters):
    fn = P.Rosenbrock(tree)
    if lam is not None:
        fn = P.L2Regularized(fn, lam)
    if box is not None:
        fn = P.UniformBox(fn, *box)
    opt = P.LegacyLBFGSOptimizer(fn, list(x0), step, m, mi, tree)
    rows = []
    for _ in range(iters):
        opt.step()
        rows.append({"f": float(opt.current_objective_value).hex(), "L": float(opt.last_step_length).hex(),
                     "iter": opt.iteration_count, "term": opt.has_terminated, "hist": opt._history_count})
        if opt.has_terminated:
            break
    return {"x0": hx(x0), "step": step, "m": m, "max_increases": mi, "lambda": lam, "box": box, "tree": tree, "rows": rows,
            "final_x": hx(opt.current_point), "final_d": hx(opt.next_step_direction), "rho": hx(opt._rho), "alpha": hx(opt._alpha)}


def live_lbfgs_trace(x0, step, m, tree, iters):
    opt = P.LiveLBFGSOptimizer(P.Rosenbrock(tree), list(x0), step, m, tree)
    rows = []
    for _ in range(iters):
        opt.step()
        rows.append({"f": float(opt.current_objective_value).hex(), "iter": opt.iteration_count, "stuck": opt.is_stuck})
    return {"x0": hx(x0), "step": step, "m": m, "rows": rows, "final_x": hx(opt.current_point),
            "final_d": hx(opt.step_direction), "rho": hx(opt.rho)}


def main_next():
    """Fixtures of the SURVEY 8f rows (second file so that golden.json stays byte-identical)."""
    g = {}
    u = [4.0 * a - 2.0 for a in P.pcg_fill(64, 5)]
    g["legacy_lbfgs"] = [legacy_trace(u[:n], 1.0, m, mi, lam, box, True, 25)
                         for n, m, mi, lam, box in ((2, 3, 0, None, None), (16, 5, 0, None, None), (34, 4, 2, 0.1, None),
                                                    (64, 3, 0, None, [-0.5, 0.8]), (16, 1, 0, 0.01, [-1.0, 0.5]))]
    g["live_lbfgs"] = [live_lbfgs_trace(u[:n], 1.0, m, True, 25) for n, m in ((2, 2), (34, 5), (64, 3))]
    # DZO_ORDER_TREE_BLOCKED above one block: dot and extended Rosenbrock of 70000 elements
    n = 70000
    a = P.pcg_fill(2 * n, 11)
    x = [4.0 * t - 2.0 for t in a[:n]]
    y = [4.0 * t - 2.0 for t in a[n:]]
    g["blocked_n70000"] = {"seed": 11, "dot": P.dot(x, y, P.BLOCKED).hex(), "rosenbrock": P.Rosenbrock(P.BLOCKED).f(x).hex()}
    # live LineSearchEvaluator
    fn = P.Rosenbrock(True)
    x = u[:34]
    gr = [0.0] * 34
    fn.g(gr, x)
    d = [-t for t in gr]
    f_old = fn.f(x)
    overlap = P.dot(gr, d, True)
    tp, tg, f_new, ir, sr = P.live_line_search_evaluate(fn, x, f_old, d, overlap, 1e-3, True, True)
    g["line_search_evaluator_n34"] = {"x": hx(x), "f_old": f_old.hex(), "overlap": overlap.hex(), "step": 1e-3,
                                      "trial_point": hx(tp), "trial_gradient": hx(tg), "f_new": f_new.hex(),
                                      "improvement_ratio": ir.hex(), "slope_ratio": sr.hex()}
    with open(os.path.join(HERE, "golden_next.json"), "w") as f:
        json.dump(g, f, indent=0, sort_keys=True)
    print("wrote golden_next.json:", {k: (len(v) if isinstance(v, list) else "ok") for k, v in g.items()})


if __name__ == "__main__":
    if "--next-only" not in sys.argv:
        main()
    main_next()
