"""Batches of medium-n BFGS problems (32 < n: a warp or a thread-block cluster per problem for the O(n) stage, blockIdx.z of
the n^2 sweeps): ms per step! call and achieved HBM GB/s.  A BFGS-type step moves 24 n^2 bytes (GEMV read + fused
update read/write), a gradient-descent step 8 n^2 (identity_matrix! write); the kinds are read back after every call
(last_step_type of the problems whose iteration_count moved), every call is timed with CUDA events on the launching stream."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    import dzopt_b200 as dz
    EF = dz.ExampleFunctions
    peak, _ = bench.load_peaks()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    cases = ((64, 16384), (128, 8192), (256, 2048), (512, 1024), (1024, 256), (2048, 64))
    for knob in ((1, 0) if "--ab" in sys.argv else (1,)):
        dz.set_tuning("warp_search", knob)
        for n, batch in cases:
            x0 = (4.0 * dz.pcg_fill(n * batch, 7) - 2.0).reshape(batch, n)
            opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
            opt.set_stream(stream.cuda_stream)
            opt.step(3)
            steps = 10
            ms_total, nbfgs, ngd, moved_total = 0.0, 0, 0, 0
            for _ in range(steps):
                it0 = opt.iteration_count.copy()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                opt.step_async(1)
                e1.record(stream)
                torch.cuda.synchronize()
                ms_total += e0.elapsed_time(e1)
                moved = (opt.iteration_count - it0) > 0
                ty = opt.last_step_type
                nbfgs += int((moved & (ty == dz.StepType.BFGSStep)).sum())
                ngd += int((moved & (ty == dz.StepType.GradientDescentStep)).sum())
                moved_total += int(moved.sum())
            ms = ms_total / steps
            nbytes = (24.0 * nbfgs + 8.0 * ngd) * n * n / steps
            gbs = nbytes / (ms * 1e-3) / 1e9
            print(json.dumps({"n": n, "batch": batch, "warp_search": knob, "ms_per_step_call": ms,
                              "problem_steps_per_s": moved_total / steps / (ms * 1e-3), "bfgs_type_steps": nbfgs, "gd_type_steps": ngd,
                              "algorithmic_bytes_per_call": nbytes, "achieved_gbs": gbs, "frac_of_peak": gbs / peak}), flush=True)
            opt.close()
    dz.set_tuning("warp_search", 1)


if __name__ == "__main__":
    main()
