#!/bin/bash
# A/B over environment switches: prints graph-step ms for each setting, two rounds (boxes drift by a few %).
# usage: tools/ab_env.sh "A=0 B=0" "A=1 B=0" ...
for r in 1 2; do
  for cfg in "$@"; do
    ms=$(env $cfg python bench.py --no-cpu-baseline --no-encoder --steps 300 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"])')
    echo "round $r  [$cfg]  ms_per_step $ms"
  done
done
