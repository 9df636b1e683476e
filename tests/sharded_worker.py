"""Row-sharded large-n BFGS (BASELINE configs[3]) -- one process per GPU.

    torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tests/sharded_worker.py \
        --size 4096 --steps 8 [--check] [--time]

--check : every rank compares its replicated vectors and its row slab of the inverse Hessian with
          the UNSHARDED CPU oracle (TREE order) bit for bit: results must not depend on the rank count.
--time  : CUDA-event time per BFGS-type step!, max over ranks; prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--tune", action="append", default=[], help="key=value for dzo_set_tuning")
    ap.add_argument("--nccl", action="store_true", help="use ncclAllGather instead of the fused peer-memory gathers")
    ap.add_argument("--delay-rank", type=int, default=-1,
                    help="--check: this rank sleeps before every step!, the others read next_step_direction right behind theirs")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import ctypes as C
    import dzopt_b200 as dz
    import oracle as orc
    from conftest import assert_bitwise
    EF = dz.ExampleFunctions
    if args.nccl:
        dz.set_tuning("sharded_variant", 1)
    for kv in args.tune:
        k, v = kv.split("=")
        dz.set_tuning(k, int(v))

    # the 128-byte ncclUniqueId travels over the host framework's own plumbing
    idbuf = C.create_string_buffer(128)
    if rank == 0:
        rc = dz.lib().dzo_nccl_get_unique_id(idbuf)
        assert rc == 0, dz.lib().dzo_last_error()
    box = [bytes(idbuf.raw)]
    dist.broadcast_object_list(box, src=0)

    n = args.size
    x0 = 4.0 * orc.pcg_fill(n, 2) - 2.0
    opt = dz.BFGSOptimizer(EF.rosenbrock_function, EF.rosenbrock_gradient_, x0, 1.0, device=local,
                           shard=(rank, world, box[0]))
    mode = {0: "single", 1: "fused-peer-memory", 2: "nccl-allgather"}[opt.gather_mode]
    r0, r1 = opt.row_range
    assert (r0, r1) == (rank * n // world, (rank + 1) * n // world)

    if args.check:
        # the oracle's whole trace FIRST: nothing but the step! itself may sit between a rank's step! and its read of
        # next_step_direction (a CPU oracle step in between used to hide a missing wait for the peers' rows)
        import time
        ref = orc.BFGS(orc.OBJ_ROSENBROCK, x0[None, :], 1.0, order=orc.TREE, nthreads=4)
        trace = []
        for it in range(args.steps):
            ref.step(1)
            trace.append((ref.point[0], ref.gradient[0], ref.direction[0], float(ref.objective[0]), float(ref.step_length[0])))
        dist.barrier()
        types = []
        for it in range(args.steps):
            if rank == args.delay_rank:
                time.sleep(0.25)
            dz.step_(opt)
            d_now = opt.next_step_direction          # immediately behind step!: every peer's rows must have landed
            rp, rg, rd, rf, rl = trace[it]
            assert_bitwise(d_now, rd, f"rank {rank} iter {it} direction (read right behind step!)")
            assert_bitwise(opt.current_point, rp, f"rank {rank} iter {it} point")
            assert_bitwise(opt.current_gradient, rg, f"rank {rank} iter {it} gradient")
            assert float(opt.current_objective_value[()]) == rf
            assert float(opt.last_step_length[()]) == rl
            types.append(int(opt.last_step_type[()]))
        assert dz.StepType.BFGSStep in types
        assert_bitwise(opt.inverse_hessian(), ref.inverse_hessian(0)[r0:r1], f"rank {rank} H slab")
        ok = torch.ones(1, device="cuda")
        dist.all_reduce(ok)
        if rank == 0:
            print(f"sharded check ok: n={n} ranks={world} steps={args.steps} types={types} gather={mode}")

    if args.time:
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        opt.set_stream(stream.cuda_stream)
        opt.step(3)
        dist.barrier(); torch.cuda.synchronize()
        calls0, _ = opt.step_log()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        evs[0].record(stream)
        for i in range(args.steps):          # back to back: no host sync between step! calls
            opt.step_async(1)
            evs[i + 1].record(stream)
        torch.cuda.synchronize()
        tt = torch.tensor([evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        times = tt.tolist()
        _, kinds = opt.step_log()
        types = [int(kinds[(calls0 + i) % 64]) for i in range(args.steps)]
        bf = [t for t, ty in zip(times, types) if ty == dz.StepType.BFGSStep]
        if rank == 0:
            peak = 6456.8
            try:
                peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            except Exception:
                pass
            ms = float(np.mean(bf)) if bf else float("nan")
            gbs = 24.0 * n * n / world / (ms * 1e-3) / 1e9
            print(json.dumps({"workload": f"row-sharded BFGS n={n}", "n_gpus": world, "bfgs_steps": len(bf),
                              "ms_per_bfgs_step": ms, "steps_per_s": 1e3 / ms, "per_gpu_achieved_gbs": gbs,
                              "frac_of_peak": gbs / peak, "rows_per_gpu": n // world, "gather": mode, "tune": args.tune}))
    opt.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
