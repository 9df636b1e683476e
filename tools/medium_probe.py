"""Timing of a batch of medium-n BFGS problems (32 < n: one thread-block cluster per problem for the O(n) stage, blockIdx.z of
the n^2 sweeps): ms per step! call and achieved HBM GB/s of the BFGS-type steps (24 n^2 bytes each)."""
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
    for n, batch in ((64, 4096), (256, 1024), (1024, 256), (2048, 64)):
        x0 = (4.0 * dz.pcg_fill(n * batch, 7) - 2.0).reshape(batch, n)
        opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
        opt.set_stream(stream.cuda_stream)
        opt.step(3)
        steps = 10
        it0 = opt.iteration_count.copy()
        types = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            opt.step_async(1)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        moved = int((opt.iteration_count - it0).sum())
        # upper bound of the traffic: every moved problem-step counted as BFGS-type (24 n^2 bytes)
        gbs = 24.0 * n * n * moved / steps / (ms * 1e-3) / 1e9
        print(json.dumps({"n": n, "batch": batch, "ms_per_step_call": ms, "problem_steps_per_s": moved / steps / (ms * 1e-3),
                          "gbs_if_all_bfgs_type": gbs, "frac_of_peak_upper_bound": gbs / peak}), flush=True)
        opt.close()


if __name__ == "__main__":
    main()
