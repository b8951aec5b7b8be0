#!/bin/bash
# GPU call 15: LighterGlue with the two sides on two streams (parity + timing), lazy-rescale threshold sweep of the attention kernel.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "lighter" 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/r2_ltg_streams_tests.log
timeout 300 python tools/bench_widened.py --only lighterglue 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/r2_ltg_streams_widened.log
for lz in 0 4 8 16; do
  DIMB_ATTN_LAZY=$lz timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_q15_lazy$lz.json 2>gpurun_out/r2_q15_lazy$lz.err; python - <<P
import json
d=json.load(open('gpurun_out/r2_q15_lazy$lz.json')); k=d['kernels_ms_per_step']
print('lazy $lz', round(d['value'],1), 'attn', round(k['lg.attn_self']+k['lg.attn_cross'],2))
P
done
