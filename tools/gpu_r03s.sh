#!/bin/bash
# round 2, GPU call 3S (1 GPU): batch-aware threads per sweep CTA -- parity of the medium / large paths, medium-n probe
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bfgs.py -m gpu -x -q -k "medium or large or warp or update or gemv or identity" > gpurun_out/r03s_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r03s_pytest.log
timeout 600 python tools/medium_probe.py > gpurun_out/r03s_medium.jsonl 2>&1; cut -c1-120 gpurun_out/r03s_medium.jsonl; python -c "
import json
for l in open('gpurun_out/r03s_medium.jsonl'):
    d=json.loads(l); print(d['n'], round(d['ms_per_step_call'],4), round(d['frac_of_peak'],3))"
