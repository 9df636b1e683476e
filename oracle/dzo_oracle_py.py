"""Independent pure-Python restatement of the same reference path (small cases only).

TEST INFRASTRUCTURE.  Written from the Julia sources (legacy/DZOptimization.jl,
legacy/Kernels.jl, legacy/ExampleFunctions.jl, legacy/PCG.jl) separately from
oracle/dzo_oracle.c, with Python floats (IEEE binary64, no FMA), so that a transcription
slip in either restatement shows up as a bitwise mismatch in tests/test_oracle.py.
It also generates the committed fixtures in tests/golden/ (tests/golden/make_golden.py).
"""
import math

INF = float("inf")
TREE_WIDTH = 4096
GEMV_CHUNK = 1024
RIESZ_SEG = 128
CAP = 4096
MASK64 = (1 << 64) - 1


# ------------------------------------------------------------------ legacy/PCG.jl
def pcg_fill(count, seed):
    def advance(s):  # :7-8
        return (0x5851F42D4C957F2D * s + 0x14057B7EF767814F) & MASK64

    def extract(s):  # :11-12  bitrotate(x, -k) == rotate right by k
        v = (((s >> 18) ^ s) >> 27) & 0xFFFFFFFF
        k = s >> 59
        return ((v >> k) | (v << (32 - k))) & 0xFFFFFFFF if k else v

    state = advance((0x14057B7EF767814F + seed) & MASK64)  # :16
    out = []
    for _ in range(count):
        out.append(2.3283064365386962890625E-10 * extract(state))  # :18
        state = advance(state)
    return out


# ------------------------------------------------------------------ summation orders
def _tree_combine(p):
    for bit in (4, 3, 2, 1, 0, 5, 6, 7, 8, 9, 10, 11):
        m = 1 << bit
        for v in range(TREE_WIDTH):
            if not v & m:
                p[v] = p[v] + p[v | m]
    return p[0]


TREE_BLOCK = 65536
BLOCKED = 2   # tree == BLOCKED: DZO_ORDER_TREE_BLOCKED (include/dzopt.h)


def ksum(terms, tree, owner=lambda k: k, block_terms=TREE_BLOCK):
    """Sum an iterable of terms: left-to-right (legacy/Kernels.jl:12-20), canonical tree, or -- tree == BLOCKED --
    canonical tree per block of `block_terms` consecutive terms with the block results added in ascending order."""
    if not tree:
        r = 0.0
        for t in terms:
            r += t
        return r
    if tree == BLOCKED and tree is not True:
        total, p, filled, first = 0.0, [0.0] * TREE_WIDTH, 0, True
        for t in terms:
            p[owner(filled) % TREE_WIDTH] += t
            filled += 1
            if filled == block_terms:
                b = _tree_combine(p)
                total = b if first else total + b
                first, p, filled = False, [0.0] * TREE_WIDTH, 0
        if filled or first:
            b = _tree_combine(p)
            total = b if first else total + b
        return total
    p = [0.0] * TREE_WIDTH
    for k, t in enumerate(terms):
        p[owner(k) % TREE_WIDTH] += t
    return _tree_combine(p)


def dot(v, w, tree=False):
    return ksum((a * b for a, b in zip(v, w)), tree, owner=lambda e: e // 2, block_terms=TREE_BLOCK)


def norm2(x, tree=False):
    return dot(x, x, tree)


def gemv(H, v, tree=False):
    """H[i][j]; row sums with j ascending (stand-in for LinearAlgebra.mul!)."""
    n = len(v)
    chunk = GEMV_CHUNK if tree else max(n, 1)
    out = []
    for i in range(n):
        acc = None
        for c0 in range(0, n, chunk):
            part = 0.0
            for j in range(c0, min(c0 + chunk, n)):
                part += H[i][j] * v[j]
            acc = part if acc is None else acc + part
        out.append(acc)
    return out


# ------------------------------------------------------------------ legacy/ExampleFunctions.jl
class Rosenbrock:
    """:10-24, extended over consecutive pairs [GLUE, SURVEY.md 8.0]"""

    def __init__(self, tree=False):
        self.tree = tree

    def constraint(self, x):
        return True

    def f(self, v):
        def term(k):
            x, y = v[2 * k], v[2 * k + 1]
            t1 = 1 - x
            t2 = y - x * x
            return t1 * t1 + 100 * (t2 * t2)
        return ksum((term(k) for k in range(len(v) // 2)), self.tree, block_terms=TREE_BLOCK // 2)

    def g(self, g, v):
        for k in range(len(v) // 2):
            x, y = v[2 * k], v[2 * k + 1]
            t1 = 1 - x
            t2 = y - x * x
            g[2 * k] = -2 * t1 - 400 * x * t2
            g[2 * k + 1] = 200 * t2


class Riesz:
    """:30-83 on a dim x N column-major matrix stored flat (point j at [j*dim, (j+1)*dim))."""

    def __init__(self, dim, sphere=False, tree=False):
        self.dim, self.sphere, self.tree = dim, sphere, tree

    def constraint(self, x):
        if self.sphere:  # [GLUE] normalise every column
            d = self.dim
            for j in range(len(x) // d):
                s = 0.0
                for k in range(d):
                    s += x[j * d + k] * x[j * d + k]
                inv = _div(1.0, math.sqrt(s))
                for k in range(d):
                    x[j * d + k] *= inv
        return True

    def _rsqrt_dist(self, p, i, j):
        d = self.dim
        dist_sq = 0.0
        for k in range(d):
            dist = p[i * d + k] - p[j * d + k]
            dist_sq += dist * dist
        return dist_sq

    def f(self, p):
        d = self.dim
        npts = len(p) // d
        if not self.tree:  # :34-43
            result = 0.0
            for j in range(1, npts):
                for i in range(j):
                    result += _div(1.0, math.sqrt(self._rsqrt_dist(p, i, j)))
            return result
        rows = []
        for j in range(npts):
            ej = 0.0
            for s0 in range(0, j, RIESZ_SEG):
                seg = 0.0
                for i in range(s0, min(s0 + RIESZ_SEG, j)):
                    seg += _div(1.0, math.sqrt(self._rsqrt_dist(p, i, j)))
                ej = seg if s0 == 0 else ej + seg
            rows.append(ej)
        return ksum(rows, True)

    def g(self, g, p):
        d = self.dim
        npts = len(p) // d
        seg = RIESZ_SEG if self.tree else npts
        for j in range(npts):
            acc = [0.0] * d
            for s0 in range(0, npts, seg):
                part = [0.0] * d
                for i in range(s0, min(s0 + seg, npts)):
                    if i == j:
                        continue
                    dist_sq = self._rsqrt_dist(p, i, j)
                    inv_dist = _div(1.0, math.sqrt(dist_sq))
                    inv_dist_cubed = _div(inv_dist, dist_sq)
                    for k in range(d):
                        part[k] += (p[i * d + k] - p[j * d + k]) * inv_dist_cubed
                acc = part if s0 == 0 else [a + b for a, b in zip(acc, part)]
            for k in range(d):
                g[j * d + k] = acc[k]
            if self.sphere:  # :361-374
                overlap = 0.0
                for k in range(d):
                    overlap += p[j * d + k] * g[j * d + k]
                for k in range(d):
                    g[j * d + k] -= overlap * p[j * d + k]


# ------------------------------------------------------------------ line search
class Ray:
    """LineSearchEvaluator (:12-46) with sign=+1; the BFGS functor [GLUE] with sign=-1."""

    def __init__(self, fn, x, direction, sign):
        self.fn, self.x, self.dir, self.sign = fn, x, direction, sign
        self.new = list(x)
        self.ref = list(x)

    def move(self, t):
        a = self.sign * t
        changed = False
        for i in range(len(self.x)):
            nw = self.x[i] + a * self.dir[i]
            changed |= (self.x[i] != nw)
            self.new[i] = nw
        return changed

    def __call__(self, t):  # :25-46
        self.move(t)
        if not self.fn.constraint(self.new):
            return INF
        return self.fn.f(self.new)


def find_three_point_bracket(lse, f0, t1, max_increases):  # :49-172
    if not math.isfinite(f0):
        return (0.0, f0, 0.0, f0)
    if not math.isfinite(t1) or t1 == 0.0:  # [GLUE]
        return (0.0, f0, 0.0, f0)
    if all(s == 0.0 for s in lse.dir):  # :71-85
        return (0.0, f0, 0.0, f0)
    step_size = t1
    point_changed = lse.move(step_size)
    step_is_small = False
    cap = CAP
    while not point_changed:  # :91-101
        step_size += step_size
        step_is_small = True
        point_changed = lse.move(step_size)
        cap -= 1
        if cap == 0:
            return (0.0, f0, 0.0, f0)
    is_feasible = lse.fn.constraint(lse.new)  # :104
    if step_is_small:
        if not is_feasible:
            return (0.0, f0, 0.0, f0)
        if lse.x == lse.new:  # :119
            return (0.0, f0, 0.0, f0)
    f1 = lse.fn.f(lse.new) if is_feasible else INF  # :126
    if f1 <= f0:  # :130
        lse.ref = list(lse.new)
        num_increases = 0
        cap = CAP
        while True:
            double_step_size = step_size + step_size
            num_increases += 1
            f2 = lse(double_step_size)
            cap -= 1
            if ((max_increases > 0 and num_increases >= max_increases) or not math.isfinite(f2)
                    or f2 > f1 or lse.new == lse.ref or cap == 0):
                return (step_size, f1, double_step_size, f2)
            step_size = double_step_size
            f1 = f2
            lse.ref = list(lse.new)
    else:  # :157-171
        cap = CAP
        while True:
            half_step_size = 0.5 * step_size
            f2 = lse(half_step_size)
            cap -= 1
            if f2 <= f0 or cap == 0:
                return (half_step_size, f2, step_size, f1)
            step_size = half_step_size
            f1 = f2


def quadratic_line_search(lse, f0, t1=1.0, max_increases=0):  # :191-216
    x1, f1, x2, f2 = find_three_point_bracket(lse, f0, t1, max_increases)
    xb, fb = 0.0, f0
    if f1 < fb:
        xb, fb = x1, f1
    if f2 < fb:
        xb, fb = x2, f2
    delta_1 = f0 - f1
    delta_2 = f2 - f1
    sum_deltas = delta_1 + delta_2
    if delta_1 >= 0.0 and delta_2 >= 0.0 and sum_deltas > 0.0:
        twice_delta_1 = delta_1 + delta_1
        delta_ratio = (twice_delta_1 + sum_deltas) / (sum_deltas + sum_deltas)
        xq = delta_ratio * x1
        fq = lse(xq)
        if fq < fb:
            xb, fb = xq, fq
    return xb, fb


# ------------------------------------------------------------------ BFGS (:725-994)
NullStep, GradientDescentStep, BFGSStep = 0, 1, 2


class BFGSOptimizer:
    def __init__(self, fn, x0, initial_step_length, tree=False):  # :762-810
        self.fn, self.tree = fn, tree
        n = len(x0)
        self.iteration_count = 0
        self.has_terminated = False
        self.current_point = list(x0)
        assert fn.constraint(self.current_point)
        self.current_objective_value = fn.f(self.current_point)
        assert not math.isnan(self.current_objective_value)
        self.current_gradient = [0.0] * n
        fn.g(self.current_gradient, self.current_point)
        self.delta_point = [0.0] * n
        self.delta_gradient = [0.0] * n
        self.last_step_length = initial_step_length
        self.last_step_type = NullStep
        self.H = [[1.0 if i == j else 0.0 for j in range(n)] for i in range(n)]
        self.next_step_direction = list(self.current_gradient)

    def update_inverse_hessian(self, step_length, step_direction, delta_gradient):  # :864-889
        n = len(step_direction)
        overlap = dot(step_direction, delta_gradient, self.tree)
        inv = 1.0 / overlap
        for i in range(n):
            step_direction[i] *= inv
        scratch = gemv(self.H, delta_gradient, self.tree)
        delta_norm = step_length * overlap + dot(delta_gradient, scratch, self.tree)
        for j in range(n):
            sj, tj = step_direction[j], scratch[j]
            for i in range(n):
                self.H[i][j] += (delta_norm * (step_direction[i] * sj)
                                 - (scratch[i] * sj + step_direction[i] * tj))

    def step(self):  # :891-994
        if self.has_terminated:
            return self
        fn, n = self.fn, len(self.current_point)
        point, gradient, d = self.current_point, self.current_gradient, self.next_step_direction
        f0 = self.current_objective_value
        step_length = self.last_step_length
        grad_norm = math.sqrt(norm2(gradient, self.tree))
        grad_step_length, grad_obj = quadratic_line_search(
            Ray(fn, point, gradient, -1.0), f0, _div(step_length, grad_norm))
        bfgs_norm = math.sqrt(norm2(d, self.tree))
        bfgs_step_length, bfgs_obj = quadratic_line_search(
            Ray(fn, point, d, -1.0), f0, _div(step_length, bfgs_norm))
        if bfgs_obj < f0 and not (bfgs_obj > grad_obj):
            self.current_objective_value = bfgs_obj
            self.last_step_length = bfgs_step_length * bfgs_norm
            self.last_step_type = BFGSStep
            self.iteration_count += 1
            self._move(-bfgs_step_length, d)
            self.update_inverse_hessian(-bfgs_step_length, d, self.delta_gradient)
            d[:] = gemv(self.H, gradient, self.tree)
        elif grad_obj < f0:
            self.current_objective_value = grad_obj
            self.last_step_length = grad_step_length * grad_norm
            self.last_step_type = GradientDescentStep
            self.iteration_count += 1
            self._move(-grad_step_length, gradient)
            self.H = [[1.0 if i == j else 0.0 for j in range(n)] for i in range(n)]
            d[:] = gradient
        else:
            self.has_terminated = True
        return self

    def _move(self, alpha, direction):  # :943-950 / :971-978
        n = len(self.current_point)
        point, gradient = self.current_point, self.current_gradient
        for i in range(n):
            self.delta_point[i] = -point[i]
            self.delta_gradient[i] = -gradient[i]
        step = [alpha * direction[i] for i in range(n)]  # direction may alias gradient
        for i in range(n):
            point[i] += step[i]
        assert self.fn.constraint(point)
        self.fn.g(gradient, point)
        for i in range(n):
            self.delta_point[i] += point[i]
            self.delta_gradient[i] += gradient[i]


def _div(a, b):
    """IEEE a / b (Python raises on b == 0; Julia returns Inf/NaN)."""
    if b == 0.0:
        return math.nan if (a == 0.0 or math.isnan(a)) else math.copysign(INF, a) * math.copysign(1.0, b)
    return a / b


# ------------------------------------------------------------------ gradient descent (:305-449)
class GradientDescentOptimizer:
    def __init__(self, fn, x0, initial_step_length, max_increases=0, tree=False):  # :330-374
        self.fn, self.tree, self.max_increases = fn, tree, max_increases
        n = len(x0)
        self.current_point = list(x0)
        assert fn.constraint(self.current_point)
        self.delta_point = [0.0] * n
        self.current_objective_value = fn.f(self.current_point)
        self.delta_objective_value = 0.0
        self.current_gradient = [0.0] * n
        fn.g(self.current_gradient, self.current_point)
        self.delta_gradient = [0.0] * n
        self.last_step_length = 0.0
        inv_gradient_norm = _div(1.0, math.sqrt(norm2(self.current_gradient, tree)))
        self.next_step_direction = [0.0] * n
        if math.isfinite(inv_gradient_norm):
            a = -initial_step_length * inv_gradient_norm
            self.next_step_direction = [gi * a for gi in self.current_gradient]
        self.iteration_count = 0
        self.has_terminated = (not math.isfinite(self.current_objective_value)
                               or not math.isfinite(inv_gradient_norm))

    def step(self):  # :393-449
        if self.has_terminated:
            return self
        n = len(self.current_point)
        x, d, g = self.current_point, self.next_step_direction, self.current_gradient
        step_size, objective_value = quadratic_line_search(
            Ray(self.fn, x, d, +1.0), self.current_objective_value, 1.0, self.max_increases)
        if step_size == 0.0 or not (objective_value < self.current_objective_value):
            self.has_terminated = True
            return self
        self.iteration_count += 1
        self.delta_point = list(x)
        for i in range(n):
            x[i] += step_size * d[i]
        assert self.fn.constraint(x)
        for i in range(n):
            self.delta_point[i] = x[i] - self.delta_point[i]
        step_length = math.sqrt(norm2(self.delta_point, self.tree))
        self.last_step_length = step_length
        self.delta_objective_value = objective_value - self.current_objective_value
        self.current_objective_value = objective_value
        self.delta_gradient = list(g)
        self.fn.g(g, x)
        for i in range(n):
            self.delta_gradient[i] = g[i] - self.delta_gradient[i]
        inv_gradient_norm = _div(1.0, math.sqrt(norm2(g, self.tree)))
        if not math.isfinite(inv_gradient_norm):
            self.has_terminated = True
            return self
        a = -step_length * inv_gradient_norm
        for i in range(n):
            d[i] = a * g[i]
        return self


# ------------------------------------------------------------------ live src/ExampleFunctions.jl (pairwise radial)
def _fma(a, b, c):
    """muladd(a, b, c) as ONE rounding (exact rational arithmetic, then round to nearest even)."""
    from fractions import Fraction
    if not (math.isfinite(a) and math.isfinite(b) and math.isfinite(c)):
        return a * b + c
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def lj_energy(r2):  # :16-27
    inv_r2 = 1.0 / r2
    inv_r4 = inv_r2 * inv_r2
    inv_r6 = inv_r4 * inv_r2
    return 4.0 * _fma(inv_r6, inv_r6, -inv_r6)


def lj_first_derivative(r2):  # :30-47
    inv_r2 = 1.0 / r2
    inv_r4 = inv_r2 * inv_r2
    inv_r6 = inv_r4 * inv_r2
    inv_r8 = inv_r4 * inv_r4
    return -12.0 * _fma(inv_r8, inv_r6 + inv_r6, -inv_r8)


def lj_second_derivative(r2):  # :50-72
    inv_r2 = 1.0 / r2
    inv_r4 = inv_r2 * inv_r2
    inv_r8 = inv_r4 * inv_r4
    inv_r10 = inv_r8 * inv_r2
    return 48.0 * _fma(3.5, inv_r8 * inv_r8, -inv_r10)


def _segments(n, tree):
    seg = RIESZ_SEG if tree else n
    return [(s0, min(s0 + seg, n)) for s0 in range(0, n, seg)]


def pairwise_point_energies(x, y, z, tree=False):  # kernel :117-149
    n = len(x)
    out = []
    for i in range(n):
        acc = None
        for s0, s1 in _segments(n, tree):
            e = 0.0
            for j in range(s0, s1):
                dx, dy, dz = x[i] - x[j], y[i] - y[j], z[i] - z[j]
                r2 = dx * dx + dy * dy + dz * dz
                e += 0.0 if i == j else lj_energy(r2)
            acc = e if acc is None else acc + e
        out.append(0.5 * acc)
    return out


def pairwise_energy(x, y, z, tree=False):  # :152-173, sum through the canonical tree [GLUE]
    return ksum(pairwise_point_energies(x, y, z, tree), True)


def pairwise_gradient(x, y, z, tree=False):  # kernel :224-262
    n = len(x)
    g = ([], [], [])
    for i in range(n):
        acc = None
        for s0, s1 in _segments(n, tree):
            ax = ay = az = 0.0
            for j in range(s0, s1):
                dx, dy, dz = x[i] - x[j], y[i] - y[j], z[i] - z[j]
                r2 = dx * dx + dy * dy + dz * dz
                f = 0.0 if i == j else lj_first_derivative(r2)
                ax += f * dx; ay += f * dy; az += f * dz
            acc = (ax, ay, az) if acc is None else (acc[0] + ax, acc[1] + ay, acc[2] + az)
        for k in range(3):
            g[k].append(acc[k] + acc[k])
    return g


def pairwise_hvp(x, y, z, u, v, w, tree=False):  # kernel :367-424
    n = len(x)
    p = ([], [], [])
    for i in range(n):
        acc = None
        for s0, s1 in _segments(n, tree):
            ax = ay = az = 0.0
            for j in range(s0, s1):
                dx, dy, dz = x[i] - x[j], y[i] - y[j], z[i] - z[j]
                du, dv, dw = u[i] - u[j], v[i] - v[j], w[i] - w[j]
                r2 = dx * dx + dy * dy + dz * dz
                f = 0.0 if i == j else lj_first_derivative(r2)
                s = 0.0 if i == j else lj_second_derivative(r2)
                overlap = dx * du + dy * dv + dz * dw
                os_ = overlap * s
                gg = os_ + os_
                ax += f * du + gg * dx; ay += f * dv + gg * dy; az += f * dw + gg * dz
            acc = (ax, ay, az) if acc is None else (acc[0] + ax, acc[1] + ay, acc[2] + az)
        for k in range(3):
            p[k].append(acc[k] + acc[k])
    return p


# ------------------------------------------------------------------ live src/DZOptimization.jl: L-BFGS
def _isequal(a, b):  # Julia isequal on floats
    if a == b:
        return math.copysign(1.0, a) == math.copysign(1.0, b)
    return math.isnan(a) and math.isnan(b)


class LiveLBFGSOptimizer:
    """src/DZOptimization.jl:321-509 with take_backtracking_step! (:107-154); history index 0 = newest."""

    def __init__(self, fn, x0, initial_step_length, history_length, tree=True):  # :347-427
        self.fn, self.tree, self.m = fn, tree, history_length
        self.current_point = list(x0)
        assert fn.constraint(self.current_point)
        self.current_objective_value = fn.f(self.current_point)
        n = len(x0)
        self.current_gradient = [0.0] * n
        fn.g(self.current_gradient, self.current_point)
        self.delta_point = [0.0] * n
        self.delta_gradient = [0.0] * n
        self.delta_objective_value = 0.0
        assert initial_step_length > 0.0
        gnorm = math.sqrt(norm2(self.current_gradient, tree))
        self.is_stuck = gnorm == 0.0
        if self.is_stuck:
            self.step_direction = [0.0] * n
        else:
            c = -initial_step_length / gnorm
            self.step_direction = [gi * c for gi in self.current_gradient]
        self.iteration_count = 0
        self.s, self.y, self.alpha, self.rho = [], [], [], []

    def _direction(self):  # :430-451
        d = list(self.current_gradient)
        k = len(self.s)
        for i in range(k):
            self.alpha[i] = dot(self.s[i], d, self.tree) / self.rho[i]
            a = -self.alpha[i]
            d = [di + a * yi for di, yi in zip(d, self.y[i])]
        if k:
            c = -self.rho[0] / dot(self.y[0], self.y[0], self.tree)
            d = [di * c for di in d]
        for i in reversed(range(k)):
            beta = dot(self.y[i], d, self.tree) / self.rho[i]
            a = -(self.alpha[i] + beta)
            d = [di + a * si for di, si in zip(d, self.s[i])]
        self.step_direction = d

    def step(self):  # :454-509
        if self.is_stuck:
            return self
        if self.iteration_count > 0:
            self._direction()
        # take_backtracking_step!  :107-154
        x, d = self.current_point, self.step_direction
        self.delta_point = list(x)
        step = 1.0
        while True:
            for k in range(len(x)):
                x[k] += step * d[k]
            if all(_isequal(a, b) for a, b in zip(x, self.delta_point)):
                self.is_stuck = True
                return self
            if self.fn.constraint(x):
                nxt = self.fn.f(x)
                if nxt < self.current_objective_value:
                    self.delta_objective_value = nxt - self.current_objective_value
                    self.current_objective_value = nxt
                    self.delta_point = [1.0 * a + (-1.0) * b for a, b in zip(x, self.delta_point)]
                    break
            x[:] = self.delta_point
            step *= 0.5
        self.delta_gradient = list(self.current_gradient)
        self.fn.g(self.current_gradient, x)
        self.delta_gradient = [1.0 * a + (-1.0) * b for a, b in zip(self.current_gradient, self.delta_gradient)]
        if len(self.s) >= self.m:
            self.s.pop(); self.y.pop()
        self.s.insert(0, list(self.delta_point)); self.y.insert(0, list(self.delta_gradient))
        if len(self.alpha) < self.m:
            self.alpha.append(0.0)
        if len(self.rho) >= self.m:
            self.rho.pop()
        self.rho.insert(0, dot(self.delta_point, self.delta_gradient, self.tree))
        self.iteration_count += 1
        return self


def _julia_min(a, b):
    if math.isnan(a):
        return a
    if math.isnan(b):
        return b
    if a == b:
        return a if math.copysign(1.0, a) < 0 else b
    return a if a < b else b


class LiveAdGDOptimizer:
    """src/DZOptimization.jl:179-312 (Algorithm 1 of MM24 + backtracking safeguard)."""

    def __init__(self, fn, x0, initial_step_length, tree=True):  # :201-271
        self.fn, self.tree = fn, tree
        self.current_point = list(x0)
        assert fn.constraint(self.current_point)
        self.current_objective_value = fn.f(self.current_point)
        n = len(x0)
        self.current_gradient = [0.0] * n
        fn.g(self.current_gradient, self.current_point)
        self.delta_point = [0.0] * n
        self.delta_gradient = [0.0] * n
        self.delta_objective_value = 0.0
        assert initial_step_length > 0.0
        gnorm = math.sqrt(norm2(self.current_gradient, tree))
        self.is_stuck = gnorm == 0.0
        s0 = 0.0 if self.is_stuck else initial_step_length / gnorm
        self.current_step_size = self.previous_step_size = s0
        self.iteration_count = 0

    def step(self):  # :274-312
        if self.is_stuck:
            return self
        previous, current = self.previous_step_size, self.current_step_size
        nxt = current
        if self.iteration_count > 0:
            theta = current / previous
            nxt *= math.sqrt(1.0 + theta)
            dgn = math.sqrt(norm2(self.delta_gradient, self.tree))
            if dgn != 0.0:
                inv_L = math.sqrt(norm2(self.delta_point, self.tree)) / dgn
                nxt = _julia_min(nxt, math.sqrt(0.5) * inv_L)
        self.previous_step_size, self.current_step_size = current, nxt
        x, g = self.current_point, self.current_gradient
        self.delta_point = list(x)
        step = -nxt
        while True:  # take_backtracking_step!  :107-154
            for k in range(len(x)):
                x[k] += step * g[k]
            if all(_isequal(a, b) for a, b in zip(x, self.delta_point)):
                self.is_stuck = True
                return self
            if self.fn.constraint(x):
                val = self.fn.f(x)
                if val < self.current_objective_value:
                    self.delta_objective_value = val - self.current_objective_value
                    self.current_objective_value = val
                    self.delta_point = [1.0 * a + (-1.0) * b for a, b in zip(x, self.delta_point)]
                    break
            x[:] = self.delta_point
            step *= 0.5
        self.delta_gradient = list(g)
        self.fn.g(g, x)
        self.delta_gradient = [1.0 * a + (-1.0) * b for a, b in zip(g, self.delta_gradient)]
        self.iteration_count += 1
        return self


# ------------------------------------------------------------------ legacy decorators (:222-296)
class L2Regularized:
    """L2RegularizationWrapper (:228-234) + L2GradientWrapper (:237-251) around an objective object."""

    def __init__(self, fn, lam):
        self.fn, self.lam, self.tree = fn, lam, fn.tree

    def constraint(self, x):
        return self.fn.constraint(x)

    def f(self, x):
        return self.fn.f(x) + self.lam * norm2(x, self.tree)

    def g(self, g, x):
        self.fn.g(g, x)
        a = self.lam + self.lam
        for i in range(len(g)):
            g[i] += a * x[i]


class UniformBox:
    """UniformBoxConstraint (:257-272) + UniformBoxGradientWrapper (:275-296) around an objective object."""

    def __init__(self, fn, lower_bound, upper_bound):
        self.fn, self.lo, self.hi, self.tree = fn, lower_bound, upper_bound, fn.tree

    def constraint(self, x):
        ok = self.fn.constraint(x)
        for i in range(len(x)):
            v = x[i]
            x[i] = self.hi if v > self.hi else (self.lo if v < self.lo else v)  # Base.clamp
        return ok

    def f(self, x):
        return self.fn.f(x)

    def g(self, g, x):
        self.fn.g(g, x)
        for i in range(len(g)):
            if (x[i] <= self.lo and g[i] >= 0.0) or (x[i] >= self.hi and g[i] <= 0.0):
                g[i] = 0.0


# ------------------------------------------------------------------ legacy L-BFGS (:458-695)
class LegacyLBFGSOptimizer:
    """legacy/DZOptimization.jl:458-695: cyclic history columns, gradient retry, two-loop as written."""

    def __init__(self, fn, x0, initial_step_length, history_length, max_increases=0, tree=True):  # :489-548
        self.fn, self.tree, self.max_increases = fn, tree, max_increases
        n = len(x0)
        self.current_point = list(x0)
        assert fn.constraint(self.current_point)
        self.delta_point = [0.0] * n
        self.current_objective_value = fn.f(self.current_point)
        self.delta_objective_value = 0.0
        self.current_gradient = [0.0] * n
        fn.g(self.current_gradient, self.current_point)
        self.delta_gradient = [0.0] * n
        self.last_step_length = 0.0
        inv_gradient_norm = _div(1.0, math.sqrt(norm2(self.current_gradient, tree)))
        self.next_step_direction = [0.0] * n
        if math.isfinite(inv_gradient_norm):
            a = -initial_step_length * inv_gradient_norm
            self.next_step_direction = [gi * a for gi in self.current_gradient]
        self.iteration_count = 0
        self.has_terminated = (not math.isfinite(self.current_objective_value)
                               or not math.isfinite(inv_gradient_norm))
        assert history_length > 0
        m = history_length
        self._alpha = [0.0] * m
        self._rho = [0.0] * m
        self._dp_hist = [[0.0] * n for _ in range(m)]   # column c of _delta_point_history
        self._dg_hist = [[0.0] * n for _ in range(m)]
        self._history_count = 0

    def _search(self):
        f0 = self.current_objective_value
        t, fv = quadratic_line_search(Ray(self.fn, self.current_point, self.next_step_direction, +1.0),
                                      f0, 1.0, self.max_increases)
        return t, fv, not (t == 0.0 or not (fv < f0))

    def step(self):  # :565-695
        if self.has_terminated:
            return self
        x, g = self.current_point, self.current_gradient
        n, m, tree = len(x), len(self._alpha), self.tree
        step_size, objective_value, ok = self._search()
        if not ok:
            a = -self.last_step_length * _div(1.0, math.sqrt(norm2(g, tree)))
            self.next_step_direction = [gi * a for gi in g]
            step_size, objective_value, ok = self._search()
            if not ok:
                self.has_terminated = True
                return self
            self._history_count = 0
        self.iteration_count += 1
        d = self.next_step_direction
        old = list(x)
        for i in range(n):
            x[i] += step_size * d[i]
        assert self.fn.constraint(x)
        self.delta_point = [x[i] - old[i] for i in range(n)]
        step_length = math.sqrt(norm2(self.delta_point, tree))
        self.last_step_length = step_length
        self.delta_objective_value = objective_value - self.current_objective_value
        self.current_objective_value = objective_value
        gold = list(g)
        self.fn.g(g, x)
        self.delta_gradient = [g[i] - gold[i] for i in range(n)]
        inv_gradient_norm = _div(1.0, math.sqrt(norm2(g, tree)))
        if not math.isfinite(inv_gradient_norm):
            self.has_terminated = True
            return self
        c = (self.iteration_count - 1) % m
        self._dp_hist[c] = list(self.delta_point)
        self._dg_hist[c] = list(self.delta_gradient)
        delta_overlap = dot(self.delta_point, self.delta_gradient, tree)
        self._rho[c] = _div(1.0, delta_overlap)
        hist_count = min(self._history_count + 1, m)
        hist_end = self.iteration_count
        hist_begin = hist_end - hist_count + 1
        self._history_count = hist_count
        d = list(g)
        for it in range(hist_end, hist_begin - 1, -1):
            c = (it - 1) % m
            alpha = self._rho[c] * dot(d, self._dp_hist[c], tree)
            self._alpha[c] = alpha
            d = [di + alpha * yi for di, yi in zip(d, self._dg_hist[c])]
        gamma = _div(delta_overlap, norm2(self.delta_gradient, tree))
        d = [di * gamma for di in d]
        for it in range(hist_begin, hist_end + 1):
            c = (it - 1) % m
            beta = self._alpha[c] - self._rho[c] * dot(d, self._dg_hist[c], tree)
            d = [di + beta * si for di, si in zip(d, self._dp_hist[c])]
        d = [-di for di in d]
        gradient_overlap = dot(d, g, tree)
        if not math.isfinite(gradient_overlap):
            self.has_terminated = True
        elif gradient_overlap >= 0.0:
            a = -step_length * inv_gradient_norm
            d = [a * gi for gi in g]
        self.next_step_direction = d
        return self


# ------------------------------------------------------------------ live LineSearchEvaluator (src/DZOptimization.jl:12-92)
def live_line_search_evaluate(fn, x, f_old, direction, overlap, step_size, compute_gradient, tree=True):
    """(lse)(step_size, compute_gradient) :66-92 -> (trial_point, trial_gradient | None, f_new, improvement_ratio, slope_ratio | None)"""
    trial = list(x)
    for i in range(len(trial)):
        trial[i] += step_size * direction[i]
    if not fn.constraint(trial):
        return trial, None, INF, -INF, INF
    f_new = fn.f(trial)
    improvement_ratio = _div(f_new - f_old, step_size * overlap)
    if not compute_gradient:
        return trial, None, f_new, improvement_ratio, None
    tg = [0.0] * len(trial)
    fn.g(tg, trial)
    slope_ratio = _div(dot(tg, direction, tree), overlap)
    return trial, tg, f_new, improvement_ratio, slope_ratio
