#!/bin/bash
# GPU call 24: ALIKED tensor-core convolutions with compile-time im2col indices (parity + A/B + tiled mode), then the full suite and the
# sanitizer passes on the final build.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py -m gpu -q -x -s -k "aliked or cfg3_aliked" 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/r2_altc2_tests.log
for a in 0 1; do
  DIMB_AL_TC=$a timeout 300 python tools/bench_widened.py --only aliked 2>&1 | tail -1 | cut -c1-900 | tee gpurun_out/r2_altc2_widened_$a.log
done
timeout 300 python bench.py --mode tiled > gpurun_out/r2_mode_tiled_final.json 2> gpurun_out/r2_mode_tiled_final.err; tail -c 200 gpurun_out/r2_mode_tiled_final.err; head -c 900 gpurun_out/r2_mode_tiled_final.json
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -12 | cut -c1-400 | tee gpurun_out/r2_tests_final2.log
timeout 600 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/r2_compute_sanitizer_racecheck_smoke.log 2>&1; tail -3 gpurun_out/r2_compute_sanitizer_racecheck_smoke.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "aliked_golden" > gpurun_out/r2_compute_sanitizer_aliked.log 2>&1; tail -4 gpurun_out/r2_compute_sanitizer_aliked.log
timeout 240 ncu --set full --clock-control none --kernel-name-base demangled -k regex:al_conv3x3_tc_kernel -c 4 -o gpurun_out/r2_prof_altc2 -f python tools/bench_widened.py --only aliked > gpurun_out/ncu_altc2.log 2>&1; tail -2 gpurun_out/ncu_altc2.log
