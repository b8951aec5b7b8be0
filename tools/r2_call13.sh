#!/bin/bash
# GPU call 13: LighterGlue with the attention on the tensor cores (attn_hd128.cuh): trained-weights golden suite, timing.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "lighter" 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/r2_ltg_tc_tests.log
timeout 300 python tools/bench_widened.py --only lighterglue 2>&1 | tail -2 | cut -c1-600 | tee gpurun_out/r2_ltg_tc_widened.log
DIMB_TC=0 timeout 300 python tools/bench_widened.py --only lighterglue 2>&1 | tail -2 | cut -c1-600 | tee gpurun_out/r2_ltg_simt_widened.log
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gx_attn_tc_kernel -s 4 -c 1 -o gpurun_out/r2_prof_attn_hd128 -f python tools/bench_widened.py --only lighterglue > gpurun_out/ncu_attn_hd128.log 2>&1; tail -2 gpurun_out/ncu_attn_hd128.log
