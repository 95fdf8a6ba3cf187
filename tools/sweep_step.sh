#!/bin/bash
# Stem step (conv1 + conv2 forward + backward, inputs in HBM, CUDA graphs) over batch sizes and forward / backward forms:
# one JSON line per run -> stdout.   usage: tools/sweep_step.sh > profiles/r2_step_batch_sweep.jsonl
for b in 16 32 64 128 256; do
  for f in fused split; do
    for m in chained split; do
      python bench.py --batch $b --no-cpu-baseline --no-encoder --steps 100 --stem-forward $f --stem-backward $m 2>/dev/null | \
        python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'batch': $b, 'forward': '$f', 'backward': '$m', 'ms_per_step': d['ms_per_step'], 'windows_per_s': d['value'],
                  'step_roofline_frac': d['step_roofline_frac'], 'kernels_per_step': d['kernels_per_step']}))"
    done
  done
done
