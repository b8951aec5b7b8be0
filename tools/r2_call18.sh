#!/bin/bash
# GPU call 18/19: attention v7 (v5 with the P V-retired barriers - two per query tile, alternating - consumed lazily unless O is rescaled): parity + A/B.
set -x
mkdir -p gpurun_out
DIMB_ATTN=7 timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py tests/test_fast_mode.py -m gpu -q -x -k "lightglue or lg or pipe or superglue or cfg2 or chain or fast" 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_attn7_tests.log
for rep in 1 2 3; do
  for v in 5 7; do
    DIMB_ATTN=$v timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_q18_attn${v}_$rep.json 2>gpurun_out/r2_q18_attn${v}_$rep.err; python - <<P
import json
d=json.load(open('gpurun_out/r2_q18_attn${v}_$rep.json')); k=d['kernels_ms_per_step']
print('attn=$v', round(d['value'],1), 'attn ms', round(k['lg.attn_self']+k['lg.attn_cross'],2))
P
  done
done
DIMB_ATTN=7 timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:lg_attn5_kernel -s 2 -c 1 -o gpurun_out/r2_prof_attn7 -f python bench.py --quick --pairs 8 --steps 1 --warmup 3 > gpurun_out/ncu_attn7.log 2>&1; tail -2 gpurun_out/ncu_attn7.log
