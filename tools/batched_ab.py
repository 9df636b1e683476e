"""A/B of the batched BFGS step kernels inside ONE gpurun call (box-to-box variation is larger than the effects).

    python tools/batched_ab.py [--batch 1000000] [--steps 20] [--warmup 5] [--configs v0l1p2,v0l0p2,v2l0p2,...]

A config is l<batched_lazy>p<batched_prefetch>.  Prints one JSON line per config: ms per step!,
the step kinds counted on the device, the algorithmic bytes of that mix and the fraction of the HBM copy peak."""
import argparse
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1_000_000)
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--configs", default="l0p3,l1p2,l1p3,l1p4")
    args = ap.parse_args()
    import torch
    import bench
    import dzopt_b200 as dz
    EF = dz.ExampleFunctions
    peak, _ = bench.load_peaks()
    if os.environ.get("DZO_BATCHED_TILE"):
        dz.set_tuning("batched_tile", int(os.environ["DZO_BATCHED_TILE"]))
    n = args.n
    x0 = (4.0 * dz.pcg_fill(args.batch * n, 2024) - 2.0).reshape(args.batch, n)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    for rep in range(args.repeat):
        for cfg in args.configs.split(","):
            m = re.fullmatch(r"l(\d)p(\d+)", cfg)
            lz, pf = (int(g) for g in m.groups())
            v = 0
            dz.set_tuning("batched_lazy", lz); dz.set_tuning("batched_prefetch", pf)
            opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, batched=True)
            opt.set_stream(stream.cuda_stream)
            opt.step(args.warmup)
            it0, done0 = opt.iteration_count.copy(), opt.has_converged.copy()
            kinds0 = opt.step_kind_counts(reset=True) if v == 0 else None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(args.steps):
                opt.step_async(1)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            moved = float((opt.iteration_count - it0).sum() + (opt.has_converged & ~done0).sum())
            out = {"config": cfg, "rep": rep, "ms_per_step": ms, "problem_steps_per_s": moved / args.steps / (ms * 1e-3)}
            if v == 0:
                kinds = opt.step_kind_counts()
                b = bench.batched_bytes(kinds, n, lazy=bool(lz))
                out.update({"kinds": kinds, "algorithmic_bytes_per_step": b / args.steps,
                            "achieved_gbs": b / args.steps / (ms * 1e-3) / 1e9, "frac": b / args.steps / (ms * 1e-3) / 1e9 / peak})
            print(json.dumps(out), flush=True)
            opt.close()


if __name__ == "__main__":
    main()
