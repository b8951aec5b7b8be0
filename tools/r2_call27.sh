#!/bin/bash
# GPU call 27: pairs per step (37 / 74 / 111): launch-overhead amortisation vs. working-set effects.
set -x
mkdir -p gpurun_out
for p in 37 74 111 37 74; do
  timeout 300 python bench.py --quick --pairs $p --steps 8 --warmup 3 > gpurun_out/r2_q27_p${p}.json 2>gpurun_out/r2_q27_p${p}.err; python - <<P
import json
d=json.load(open('gpurun_out/r2_q27_p${p}.json'))
print('pairs $p', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2))
P
done
