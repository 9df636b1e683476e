#!/bin/bash
for cfg in "use_graph=1" "use_graph=0" "use_graph=1" "use_graph=0"; do
  python bench.py --steps 5 --warmup 3 --skip-cpu --tune $cfg 2>&1 | python tools/brief.py "[$cfg]" | sed 's/value=.*large:/large:/'
done
