#!/bin/bash
# GPU call 17: 32-wide K stages (gemm.cuh CONV 3) for the 128 x 256 LightGlue tiles: parity + A/B; the selectable-variant tests.
set -x
mkdir -p gpurun_out
DIMB_K32=1 timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py tests/test_fast_mode.py -m gpu -q -x -k "lightglue or lg or pipe or cfg2 or chain or fast" 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_k32_tests.log
for rep in 1 2; do
  for k in 0 1; do
    DIMB_K32=$k timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_q17_k32_${k}_$rep.json 2>gpurun_out/r2_q17_k32_${k}_$rep.err; python - <<P
import json
d=json.load(open('gpurun_out/r2_q17_k32_${k}_$rep.json')); k=d['kernels_ms_per_step']
print('k32=$k', round(d['value'],1), 'qk', k['lg.qk'], 'ffn0', k['lg.ffn0'])
P
  done
done
