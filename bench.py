#!/usr/bin/env python
"""bench.py -- batched BFGS step! throughput (BASELINE.json configs[1]) + n=16384 step! roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (N=1): 1,000,000 independent extended-Rosenbrock problems, n=16, x0 = 4u-2 with u from
the reference's PCG (legacy/PCG.jl, seed 2024+rank), initial step 1.0.  One "step" = one step!
call on every problem of the batch = ONE launch of bfgs_batched_step_kernel.  For N>1 every rank
holds its own 1M problems (weak scaling, no collective on the data path).

Numbers on the JSON line
  value     active problem-steps/s, state resident in HBM, CUDA events on the launching stream
  e2e       the README loop through the public API with HOST buffers: constructor from pinned
            host x0 (H2D) + K x [step!; read has_converged[] and current_objective_value[] (D2H)]
  roofline  bfgs_batched_step_kernel against the measured HBM copy bandwidth
  large_n   n=16384 single-problem step! (configs[2]): ms per BFGS-type step, achieved GB/s of
            the 24 n^2-byte step and of its two n^2 kernels against the same peak
  cpu_baseline  the CPU oracle (a port of the reference: the reference itself is commented-out
            Julia and no Julia exists here) on all host threads, bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SMALL = 16
BATCH = 1_000_000
LARGE_N = 16384
METRIC = "batched BFGS problem-steps/s (1M x n=16)"
UNIT = "problem-steps/s"
# algorithmic bytes of one batched problem-step: read x,g,d,H + f,L,iter,term;
# write x,g,d,dx,dg,H + f,L,iter,type
BYTES_PER_PROBLEM_STEP = (3 * N_SMALL * 8 + N_SMALL * N_SMALL * 8 + 25) + (5 * N_SMALL * 8 + N_SMALL * N_SMALL * 8 + 28)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/ncu_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def oracle_mod():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.build()
    return oracle


def x0_batch(gen, batch, seed):
    """gen: anything with pcg_fill(count, seed) -- the product package in the GPU arm, the oracle in the reference arm"""
    return (4.0 * gen.pcg_fill(batch * N_SMALL, seed) - 2.0).reshape(batch, N_SMALL)


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled every
    2 ms from a thread (nvidia-smi's fastest loop is too coarse for a 30 ms region)."""

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.nvml = None
        self.max_mhz = None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices; resolve through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(hh).bus == bus:
                        h = hh
                        break
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml, self.h = pynvml, h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.nvml = None
        return self

    def _poll(self):
        nv = self.nvml
        while not self.stop:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                                  if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons")
                                  else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
            except Exception:
                pass
            time.sleep(0.002)

    def __exit__(self, *a):
        self.stop = True
        if self.nvml:
            self.t.join(timeout=1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = set()
        for _, r in self.rows:
            for b, nm in bits.items():
                if r & b:
                    reasons.add(nm)
        return {"sm_mhz": float(np.median([c for c, _ in self.rows])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ----------------------------------------------------------------------------- CPU legs
def cpu_leg(orc, steps, warmup, sample_batch):
    """Oracle (port of the reference path) on every host thread; bounded sample of the workload."""
    threads = os.cpu_count() or 1
    x0 = x0_batch(orc, sample_batch, 2024)
    ref = orc.BFGS(orc.OBJ_ROSENBROCK, x0, 1.0, order=orc.SEQ, nthreads=threads)
    ref.step(warmup)
    active0 = ref.count_active()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        done += ref.count_active()
        ref.step(1)
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sample_batch} of the {BATCH} problems (same PCG stream), {steps} step! calls after {warmup} warm-up, "
                      f"{active0} active at start; oracle/dzo_oracle.c, OpenMP over problems",
            "seconds": dt, "ms_per_step": 1e3 * dt / steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc = oracle_mod()
    sample = 100_000
    leg = cpu_leg(orc, args.steps, args.warmup, sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": leg["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batched BFGS 1,000,000 x n=16 extended Rosenbrock (BASELINE configs[1])",
                   "sample_batch": sample, "note": "reference = CPU oracle port (Julia reference is dead code; no Julia here)"},
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import dzopt_b200 as dz
    EF = dz.ExampleFunctions
    for kv in args.tune:
        k, v = kv.split("=")
        dz.set_tuning(k, int(v))
    # the checker: only the cpu_baseline legs below touch it (inputs come from the product's own dz.pcg_fill)
    orc = oracle_mod() if (rank == 0 and world == 1 and not args.skip_cpu) else None
    peak, peak_src = load_peaks()
    K, W = args.steps, args.warmup

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if not distributed:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if not distributed:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    x0 = x0_batch(dz, BATCH, 2024 + rank)
    x0_pinned = torch.from_numpy(x0).pin_memory()
    x0_host = x0_pinned.numpy()
    stream = torch.cuda.Stream()      # a real (non-default) stream: handle 0 would mean "the handle's own stream"
    torch.cuda.set_stream(stream)

    # ---- device-resident throughput: K step! calls, one kernel launch each
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0_host, 1.0, batched=True,
                           device=local_rank)
    opt.set_stream(stream.cuda_stream)
    opt.step(W)
    # per-step active counts are read AFTER the timed region from iteration counts, not inside it
    it0 = opt.iteration_count.copy()
    done0 = opt.has_converged.copy()
    barrier()
    with ClockSampler(local_rank) as clk:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(K):
            opt.step_async(1)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
    ms = max_over_ranks(ms)
    it1 = opt.iteration_count
    done1 = opt.has_converged
    # a problem active at the start of a step either moved (iteration +1) or terminated in that step
    problem_steps = float((it1 - it0).sum() + (done1 & ~done0).sum())
    total_problem_steps = sum_over_ranks(problem_steps)
    value = total_problem_steps / (ms * 1e-3)
    kernel_ms = ms / K
    achieved = BYTES_PER_PROBLEM_STEP * problem_steps / K / (kernel_ms * 1e-3) / 1e9
    active_frac = problem_steps / (K * BATCH)
    opt.close()

    # ---- k step! calls fused into ONE launch ("report also with k fused steps per launch")
    fused = None
    if rank == 0 and not args.skip_large:
        kf = 50
        f_opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0_host, 1.0, batched=True,
                                 device=local_rank)
        f_opt.set_stream(stream.cuda_stream)
        f_opt.step(W)
        f_it0, f_done0 = f_opt.iteration_count.copy(), f_opt.has_converged.copy()
        fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fe0.record(stream); f_opt.step_async(kf); fe1.record(stream)
        torch.cuda.synchronize()
        f_ms = fe0.elapsed_time(fe1)
        f_steps = float((f_opt.iteration_count - f_it0).sum() + (f_opt.has_converged & ~f_done0).sum())
        fused = {"k": kf, "launches": 1, "ms": f_ms, "problem_steps": f_steps, "problem_steps_per_s": f_steps / (f_ms * 1e-3),
                 "still_active_after": int(f_opt.count_active())}
        f_opt.close()

    # ---- end to end through the public API with host buffers
    # untimed warm-up pass of the same code path: fills the device memory pool and the recycled page-locked
    # field buffers (page-locking costs 10-40 ms on this virtualised host), like the W warm-up steps do for kernels
    w_opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0_host, 1.0, batched=True,
                             device=local_rank)
    w_opt.reuse_host_buffers(True)
    w_opt.step(1)
    _ = w_opt.has_converged, w_opt.current_objective_value
    w_opt.close()
    barrier()
    barrier()
    t1 = time.perf_counter()
    e2 = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0_host, 1.0, batched=True,
                          device=local_rank)           # H2D of x0 inside the timed region
    e2.reuse_host_buffers(True)                         # cached page-locked field arrays; has_converged / objective become
                                                        # zero-copy mirrors the step kernel writes over PCIe while it runs
    e2.step(W)                                          # same starting state as the device-timed arm
    flags = obj = None
    for _ in range(K):
        dz.step_(e2)
        flags = e2.has_converged                        # D2H, what `while !opt.has_converged[]` reads
        obj = e2.current_objective_value                # D2H
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t1
    # every step! of an active problem either moves it (iteration_count + 1) or terminates it; the
    # warm-up steps of this arm are inside its timed region and count as work too
    e2e_steps_total = float(e2.iteration_count.sum() + flags.sum())
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = sum_over_ranks(e2e_steps_total) / e2e_s
    h2d = x0.nbytes / (K + W)
    d2h = (flags.nbytes + obj.nbytes)
    e2.close()

    # ---- large-n step! (configs[2]) and its two n^2 kernels, rank 0 only at N=1
    large = None
    if rank == 0 and not args.skip_large:
        large = bench_large(dz, orc, torch, stream, peak, local_rank, cpu=(world == 1 and not args.skip_cpu))

    riesz = None
    if rank == 0 and not args.skip_large:
        riesz = bench_riesz(dz, orc, torch, stream, local_rank, cpu=(world == 1 and not args.skip_cpu))
    readme = bench_readme(dz, orc) if (rank == 0 and world == 1 and not args.skip_cpu) else None
    lbfgs = bench_lbfgs(dz, orc, torch, stream, cpu=(world == 1 and not args.skip_cpu)) if (rank == 0 and not args.skip_large) else None
    pairwise = bench_pairwise(dz, orc, torch, cpu=(world == 1 and not args.skip_cpu)) if (rank == 0 and not args.skip_large) else None

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = cpu_leg(orc, K, W, args.cpu_sample)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batched BFGS 1,000,000 x n=16 extended Rosenbrock per GPU (BASELINE configs[1])",
                       "batch_per_gpu": BATCH, "n": N_SMALL, "initial_step_length": 1.0,
                       "l2": "state per GPU = 2.9 GB >> 126 MB L2 (inputs larger than L2, no flush needed)",
                       "active_fraction_in_timed_region": active_frac,
                       "parallelism": f"independent problems, {world} GPU(s), no collective"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "BFGSOptimizer(host x0) + W+K x [step!; has_converged[]; current_objective_value[]] via the C ABI; the two fields reach the host as zero-copy mirrors (dzo_bfgs_mirror_fields): the step kernel stores them into page-locked host memory, the reads only synchronise"},
            "gpu_launches": K,
            "roofline": {"bound": "hbm", "kernel": "bfgs_batched_hybrid_kernel<16>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic("bfgs_batched_hybrid_kernel<16>"),
                         "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)",
                         "algorithmic_bytes_per_launch": BYTES_PER_PROBLEM_STEP * BATCH,
                         "peak_source": peak_src, "bytes_per_problem_step": BYTES_PER_PROBLEM_STEP},
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if fused:
            line["fused_k_steps"] = fused
        if large:
            line["large_n"] = large
        if riesz:
            line["riesz_gd"] = riesz
        if readme:
            line["readme_rosenbrock_n2"] = readme
        if pairwise:
            line["pairwise_radial"] = pairwise
        if lbfgs:
            line["live_lbfgs"] = lbfgs
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def bench_large(dz, orc, torch, stream, peak, device, cpu=True):
    import ctypes as C
    EF = dz.ExampleFunctions
    n = LARGE_N
    x0 = 4.0 * dz.pcg_fill(n, 1) - 2.0
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, device=device)
    opt.set_stream(stream.cuda_stream)
    opt.step(3)
    # 20 step! calls enqueued back to back (no host sync in between), an event after each; the device-side
    # step log says afterwards which calls were BFGS-type (GEMV + fused update) and which reset H
    nsteps = 20
    calls0, _ = opt.step_log()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps + 1)]
    evs[0].record(stream)
    for i in range(nsteps):
        opt.step_async(1)
        evs[i + 1].record(stream)
    torch.cuda.synchronize()
    times = [evs[i].elapsed_time(evs[i + 1]) for i in range(nsteps)]
    calls1, kinds = opt.step_log()
    types = [int(kinds[(calls0 + i) % 64]) for i in range(nsteps)]
    opt.close()
    bf = [t for t, ty in zip(times, types) if ty == dz.StepType.BFGSStep]
    out = {"n": n, "steps_timed": len(times), "bfgs_steps": len(bf), "timing": "back-to-back launches, CUDA event after every step!"}
    if bf:
        ms = float(np.mean(bf))
        gbs = 24.0 * n * n / (ms * 1e-3) / 1e9
        out.update({"ms_per_bfgs_step": ms, "steps_per_s": 1e3 / ms, "achieved_gbs": gbs, "frac_of_peak": gbs / peak,
                    "algorithmic_bytes_per_step": 24 * n * n})
    if cpu:
        # the same step! on the host: oracle, TREE order, every host thread over rows (bit-identical to 1 thread)
        threads = os.cpu_count() or 1
        ref = orc.BFGS(orc.OBJ_ROSENBROCK, x0[None, :], 1.0, order=orc.TREE, nthreads=threads)
        ref.step(3)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); ref.step(1); ts.append((time.perf_counter() - t0, int(ref.step_type[0])))
        bfc = [t for t, ty in ts if ty == 2]
        if bfc:
            out["cpu_baseline"] = {"ms_per_bfgs_step": 1e3 * float(np.mean(bfc)), "cores": threads, "kind": "port",
                                   "sample": f"{len(bfc)} BFGS-type step! calls of the same n=16384 problem, oracle with OpenMP over rows"}
    ms = C.c_float()
    for which, name, nbytes in ((1, "gemv_kernel", 8), (2, "update_gemv_kernel", 16), (3, "identity_kernel", 8)):
        rc = dz.lib().dzo_bench_kernel(which, n, 10, 0, C.byref(ms), device)
        if rc == 0:
            g = nbytes * n * n / (ms.value * 1e-3) / 1e9
            out[name] = {"ms": ms.value, "achieved_gbs": g, "frac_of_peak": g / peak}
    return out


def bench_riesz(dz, orc, torch, stream, device, cpu=True):
    """BASELINE configs[4]: GradientDescentOptimizer on Riesz-energy points on the sphere, N=4096 (n=12288).
    FP64-pipe bound (one sqrt + one or two divisions per pair), not an HBM roofline: reported as GD steps/s
    and pair terms/s.  Inputs per SURVEY 8d: PCG seed 3 uniform in [-1,1)^3, normalised; initial step 1e-3."""
    EF = dz.ExampleFunctions
    N = 4096
    p = 2.0 * dz.pcg_fill(3 * N, 3).reshape(N, 3) - 1.0
    p = p / np.sqrt((p * p).sum(axis=1, keepdims=True))
    opt = dz.GradientDescentOptimizer(dz.SPHERE_CONSTRAINT, EF.riesz_energy, EF.riesz_gradient_,
                                      dz.QuadraticLineSearch(0), p, 1e-3, device=device)
    opt.set_stream(stream.cuda_stream)
    opt.step(3)
    k = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it0 = int(opt.iteration_count[()])
    e0.record(stream)
    opt.step_async(k)            # k step! calls inside ONE cooperative launch
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    done = int(opt.iteration_count[()]) - it0
    f = float(opt.current_objective_value[()])
    opt.close()
    out = {"N": N, "n": 3 * N, "gd_steps": done, "ms_per_gd_step": ms / max(done, 1), "gd_steps_per_s": 1e3 * done / ms,
           "objective": f}
    if cpu:
        ref = orc.GD(orc.OBJ_RIESZ, p.reshape(1, -1), 1e-3, order=orc.TREE, constraint=orc.CONSTRAINT_SPHERE, dim=3)
        ref.step(3)
        t0 = time.perf_counter(); ref.step(3); dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"ms_per_gd_step": 1e3 * dt / 3, "cores": 1, "kind": "port",
                               "sample": "3 GD step! calls of the same N=4096 problem, oracle single thread (the reference is single-threaded)"}
    return out


def bench_pairwise(dz, orc, torch, cpu=True):
    """the live package's accelerated pairwise radial kernels (Lennard-Jones), device arrays
    in / out as in the reference API.  FP64-pipe bound; reported as pair terms/s (n^2 per launch)."""
    EF = dz.ExampleFunctions
    n = 16384
    L = 1.2 * n ** (1.0 / 3.0)
    p = dz.pcg_fill(3 * n, 21).reshape(3, n) * L
    t = [torch.from_numpy(a.copy()).cuda() for a in p]
    g = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)]
    out = {"n": n, "potential": "lennard_jones"}
    for name, order in (("sequential", 0), ("tree", 1)):
        for _ in range(3):
            dz.accelerated_pairwise_radial_gradient_(*g, EF.lj_first_derivative, *t, order=order)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            dz.accelerated_pairwise_radial_gradient_(*g, EF.lj_first_derivative, *t, order=order)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out["gradient_" + name] = {"ms": ms, "pair_terms_per_s": n * n / (ms * 1e-3)}
    if cpu:
        t0 = time.perf_counter()
        orc.pairwise_gradient(p[0], p[1], p[2], orc.SEQ)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"ms": 1e3 * dt, "pair_terms_per_s": n * n / dt, "cores": os.cpu_count(), "kind": "port",
                               "sample": "one gradient of the same n=16384 cloud, oracle with OpenMP over particles"}
    return out


def bench_lbfgs(dz, orc, torch, stream, cpu=True):
    """the live package's LBFGSOptimizer (the package's own answer for large n) on
    extended Rosenbrock n = 2^20, history 10: k step! calls in ONE cooperative launch (eight 512-thread CTAs per block
    of 65536 elements, DZO_ORDER_TREE_BLOCKED, the direction vector in registers).  Algorithmic traffic per step! with
    a full history and one line-search trial: (4m + 12) n-vectors = every s_i and y_i twice (4m), g and x once, the
    direction stored once, accept pass (3 reads, 6 writes).  The step is barrier-bound (2m + 2 grid barriers), so the
    HBM fraction is reported for orientation, not as the bound."""
    EF = dz.ExampleFunctions
    n, m, k = 1 << 20, 10, 50
    x0 = 4.0 * dz.pcg_fill(n, 9) - 2.0
    opt = dz.LBFGSOptimizer(None, EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, m)
    opt.set_stream(stream.cuda_stream)
    opt.step(12)
    it0 = int(opt.iteration_count[()])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); opt.step_async(k); e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    done = int(opt.iteration_count[()]) - it0
    bytes_per_step = 8 * n * (4 * m + 12)
    clusters = C.c_int()
    dz.lib().dzo_lbfgs_info(opt._h, None, None, C.byref(clusters))
    out = {"n": n, "history_length": m, "steps": done, "ms_per_step": ms / max(done, 1), "steps_per_s": 1e3 * done / ms,
           "objective": float(opt.current_objective_value[()]), "ctas": clusters.value,
           "algorithmic_bytes_per_step": bytes_per_step,
           "achieved_gbs": bytes_per_step * done / (ms * 1e-3) / 1e9, "order": "DZO_ORDER_TREE_BLOCKED"}
    opt.close()
    if cpu:
        ref = orc.LBFGS(orc.OBJ_ROSENBROCK, x0, 1.0, m, orc.TREE_BLOCKED)
        ref.step(5)
        t0 = time.perf_counter(); ref.step(5); dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"ms_per_step": 1e3 * dt / 5, "cores": 1, "kind": "port",
                               "sample": "5 step! calls of the same problem, oracle single thread"}
    # the legacy LBFGSOptimizer (quadratic line search, cyclic history), same grid-wide machinery
    n2 = 1 << 20
    x1 = 4.0 * dz.pcg_fill(n2, 9) - 2.0
    leg = dz.LegacyLBFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, dz.QuadraticLineSearch(0), x1, 1.0, m)
    leg.set_stream(stream.cuda_stream)
    leg.step(12)
    it0 = int(leg.iteration_count[()])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); leg.step_async(k); e1.record(stream)
    torch.cuda.synchronize()
    done2 = int(leg.iteration_count[()]) - it0
    out["legacy_lbfgs"] = {"n": n2, "history_length": m, "steps": done2, "ms_per_step": e0.elapsed_time(e1) / max(done2, 1),
                           "objective": float(leg.current_objective_value[()])}
    leg.close()
    return out


def bench_readme(dz, orc):
    """BASELINE configs[0]: README Rosenbrock n=2 from rand(2), step 1.0, run to has_converged -- the
    reference's own CPU-runnable case (README.md:49-66 reports 2.8 us min / 5.6 us median per optimisation
    on an unspecified CPU).  Oracle: 1000 PCG seeds, one thread; GPU: the same 1000 problems as one batch."""
    x0 = np.stack([dz.pcg_fill(2, s) for s in range(1000)])
    t0 = time.perf_counter()
    ref = orc.BFGS(orc.OBJ_ROSENBROCK, x0, 1.0, order=orc.SEQ, nthreads=1)
    while ref.count_active():
        ref.step(16)
    cpu_us = 1e6 * (time.perf_counter() - t0) / 1000
    EF = dz.ExampleFunctions
    t0 = time.perf_counter()
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
    while opt.count_active():
        opt.step(16)
    gpu_us = 1e6 * (time.perf_counter() - t0) / 1000
    same = bool(np.array_equal(opt.current_point, ref.point))
    opt.close()
    return {"problems": 1000, "cpu_us_per_optimisation": cpu_us, "cpu_cores": 1, "cpu_kind": "port",
            "gpu_us_per_optimisation_batched": gpu_us, "bitwise_equal_to_oracle": same,
            "reference_published_us": {"min": 2.8, "median": 5.563, "hardware": "unspecified (README.md:62-63)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-large", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=50_000)
    ap.add_argument("--tune", action="append", default=[], help="key=value passed to dzo_set_tuning (A/B runs)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
