#!/bin/bash
# GPU call 26: cheaper drain of the conv1a result (ReLU + one multiplier per row instead of a select per element): parity + same-box A/B
# against the previous build (DIMB_LIB).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py -m gpu -q -x -k "superpoint or pipe or cfg2 or chain" 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2_drain_tests.log
for rep in 1 2; do
  for v in prev new; do
    if [ $v = prev ]; then export DIMB_LIB=$PWD/deep-image-matching_b200/libdimb200_prev.so; else unset DIMB_LIB; fi
    timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_q26_${v}_$rep.json 2>gpurun_out/r2_q26_${v}_$rep.err; python - <<P
import json
d=json.load(open('gpurun_out/r2_q26_${v}_$rep.json')); k=d['kernels_ms_per_step']
print('$v', round(d['value'],1), 'conv1ab', k.get('sp.conv1ab'))
P
  done
done
