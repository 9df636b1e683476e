#!/bin/bash
# round 2, GPU call 3L (1 GPU): throughput of the one-thread-per-problem batched kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/generic_probe.py 2>&1 | tee gpurun_out/r03l_generic.jsonl | tail -5
