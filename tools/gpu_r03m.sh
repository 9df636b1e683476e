#!/bin/bash
# round 2, GPU call 3M (1 GPU): batch-innermost inverse Hessians in the generic batched BFGS kernel -- parity, throughput
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gd.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r03m_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r03m_pytest.log
timeout 300 python tools/generic_probe.py 2>&1 | tee gpurun_out/r03m_generic.jsonl | tail -4
