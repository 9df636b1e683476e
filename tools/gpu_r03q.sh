#!/bin/bash
# round 2, GPU call 3Q (1 GPU): ncu of the final grid-wide L-BFGS kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/lbfgs_probe.py 1048576 10 5 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_lbfgs_kernel -s 2 -c 1 -o gpurun_out/r03q_lbfgs -f python tools/lbfgs_probe.py 1048576 10 5 > gpurun_out/r03q_ncu_lbfgs.log 2>&1; echo "ncu lbfgs rc=$?"
