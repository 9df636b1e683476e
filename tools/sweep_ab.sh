#!/bin/bash
# A/B the n^2 sweep kernels on one box: tools/sweep_ab.sh "sweep_unroll=16" "sweep_unroll=16 sweep_threads=256" ...
for cfg in "$@"; do
  args=""
  for kv in $cfg; do args="$args --tune $kv"; done
  python bench.py --steps 5 --warmup 3 --skip-cpu $args 2>&1 | python tools/brief.py "[$cfg]" | sed 's/value=.*large:/large:/'
done
