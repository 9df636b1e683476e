#!/bin/bash
# round 2, GPU call H (1 GPU): full parity suite, medium-n probe, the default bench line, launch list, ncu of the batched kernel in the timed window
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02h_pytest.log
timeout 300 python tools/medium_probe.py > gpurun_out/r02h_medium.log 2>&1; echo "probe rc=$?"
tail -6 gpurun_out/r02h_medium.log
timeout 900 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02h_bench.err
python tools/brief.py r02h < gpurun_out/r02h_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02h_ref.json 2> gpurun_out/r02h_ref.err; echo "ref rc=$?"
cut -c1-400 gpurun_out/r02h_ref.json
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/r02h_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02h_launches.csv \
    python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/r02h_ncu_launches.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:batched_hybrid -s 12 -c 1 -o gpurun_out/r02h_hybrid -f \
    python bench.py --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r02h_ncu.log 2>&1; echo "ncu rc=$?"
