#!/bin/bash
# GPU call 23: ALIKED blocks 1-2 as tensor-core im2col GEMMs (al_conv3x3_tc_kernel): parity (goldens, cfg3 tile flow, chain) + A/B.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py -m gpu -q -x -s -k "aliked or cfg3_aliked" 2>&1 | tail -14 | cut -c1-300 | tee gpurun_out/r2_altc_tests.log
for a in 0 1; do
  DIMB_AL_TC=$a timeout 300 python tools/bench_widened.py --only aliked 2>&1 | tail -1 | cut -c1-900 | tee gpurun_out/r2_altc_widened_$a.log
done
timeout 300 python bench.py --mode tiled > gpurun_out/r2_mode_tiled_altc.json 2> gpurun_out/r2_mode_tiled_altc.err; tail -c 200 gpurun_out/r2_mode_tiled_altc.err; head -c 1200 gpurun_out/r2_mode_tiled_altc.json
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -30 | cut -c1-400 | tee gpurun_out/r2_tests23.log
timeout 240 ncu --set full --clock-control none --kernel-name-base demangled -k regex:al_conv3x3_tc_kernel -c 4 -o gpurun_out/r2_prof_altc -f python tools/bench_widened.py --only aliked > gpurun_out/ncu_altc.log 2>&1; tail -2 gpurun_out/ncu_altc.log
