"""Row-sharded large-n BFGS across GPUs (one process per GPU, NCCL allgather of t and d)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, n, steps, *extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "sharded_worker.py"),
           "--size", str(n), "--steps", str(steps), "--check"] + list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded check ok" in r.stdout


def test_sharded_single_rank_equals_oracle(gpu):
    """nranks = 1 through the sharded constructor (no NCCL traffic) -- runs on the 1-GPU box."""
    _run(1, 2048, 6)


def test_sharded_two_ranks_equal_oracle(gpu):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (covered on CPU by tests/test_sharding_gloo.py)")
    _run(2, 4096, 8)


def test_sharded_direction_is_complete_right_behind_step(gpu):
    """Round-1 advisor finding: in the fused mode the peers' rows of next_step_direction are written by THEIR update
    kernels.  One rank is delayed before every step!; the others read d immediately after theirs returns and must see
    the oracle's bits (the last launch of a step! waits for every peer's flag).  Both gather modes."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, 4096, 8, "--delay-rank", "1")
    _run(2, 4096, 8, "--delay-rank", "0", "--nccl")
