#!/bin/bash
# GPU call 8: attention v4 (two softmax threads per row) parity + A/B, fused conv1a with relaxed cluster arrive A/B.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py tests/test_fast_mode.py -m gpu -q -x -k "lightglue or lg or pipe or superglue or cfg2 or chain or fast" 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_attn4_tests.log
for rep in 1 2; do
  for v in 3 4; do
    DIMB_ATTN=$v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2_ab_attn${v}_$rep.json 2>gpurun_out/r2_ab_attn${v}_$rep.err; cat gpurun_out/r2_ab_attn${v}_$rep.json; tail -c 200 gpurun_out/r2_ab_attn${v}_$rep.err
  done
done
for rep in 1 2; do
  for f in 0 1; do
    DIMB_FUSE1A=$f timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2_ab9_fuse${f}_$rep.json 2>gpurun_out/r2_ab9_fuse${f}_$rep.err; cat gpurun_out/r2_ab9_fuse${f}_$rep.json; tail -c 200 gpurun_out/r2_ab9_fuse${f}_$rep.err
  done
done
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2_tests8.log
timeout 600 python bench.py > gpurun_out/r2_bench_attn4.json 2> gpurun_out/r2_bench_attn4.err; tail -c 300 gpurun_out/r2_bench_attn4.err; head -c 600 gpurun_out/r2_bench_attn4.json
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:lg_attn4_kernel -s 2 -c 1 -o gpurun_out/r2_prof_attn4 -f python bench.py --quick --pairs 8 --steps 1 --warmup 3 > gpurun_out/ncu_attn4.log 2>&1; tail -2 gpurun_out/ncu_attn4.log
