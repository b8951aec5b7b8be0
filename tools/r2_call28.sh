#!/bin/bash
# GPU call 28: ALIKED tensor-core convolutions, third cut (16 warps, precomputed patch coordinates, prefetched patches): parity + A/B.
set -x
mkdir -p gpurun_out
DIMB_AL_TC=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py -m gpu -q -x -k "aliked or cfg3_aliked" 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2_altc3_tests.log
for a in 0 1; do
  DIMB_AL_TC=$a timeout 300 python tools/bench_widened.py --only aliked 2>&1 | tail -1 | grep -o '"ms_per_tile": [0-9.]*\|"al.block[12]": [0-9.]*' | tr '\n' ' '; echo " <- al_tc=$a"
done
