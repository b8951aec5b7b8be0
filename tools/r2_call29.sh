#!/bin/bash
# GPU call 29: conv2a on the CTA-pair kernel with eight epilogue warps (DIMB_PAIR=2): parity + A/B.
set -x
mkdir -p gpurun_out
DIMB_PAIR=2 timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py -m gpu -q -x -k "superpoint or pipe or cfg2" 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2_pair8_tests.log
for rep in 1 2; do
  for v in 1 2; do
    DIMB_PAIR=$v timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_q29_pair${v}_$rep.json 2>gpurun_out/r2_q29_pair${v}_$rep.err; python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2_q29_pair${v}_$rep.json')); k=d['kernels_ms_per_step']
    print('pair=$v', round(d['value'],1), 'conv2a', k.get('sp.conv2a'))
except Exception as e:
    print('pair=$v failed', e); print(open('gpurun_out/r2_q29_pair${v}_$rep.err').read()[-400:])
P
  done
done
