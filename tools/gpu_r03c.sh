#!/bin/bash
# round 2, GPU call 3C (1 GPU): n^2 sweeps over threads per CTA x columns in flight (fraction of the HBM copy peak)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/sweep_probe.py 16384 2>&1 | tee gpurun_out/r03c_sweeps.jsonl
