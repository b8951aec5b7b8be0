#!/bin/bash
# GPU call 7: fused conv1a with 8 producer warps (parity + A/B), ALIKED pixel-major map, SuperGlue input path, full suite.
set -x
mkdir -p gpurun_out
DIMB_FUSE1A=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "superpoint" 2>&1 | tail -5 | cut -c1-250 | tee gpurun_out/r2_fuse8_tests.log
for rep in 1 2; do
  for f in 0 1; do
    DIMB_FUSE1A=$f timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2_ab8_fuse${f}_$rep.json 2>gpurun_out/r2_ab8_fuse${f}_$rep.err; cat gpurun_out/r2_ab8_fuse${f}_$rep.json; tail -c 200 gpurun_out/r2_ab8_fuse${f}_$rep.err
  done
done
timeout 1700 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "small|tiny|tile [0-9]|passed|failed|FAILED|rror|assert" | cut -c1-300 | tee gpurun_out/r2_tests7.log | tail -25
timeout 300 python bench.py --mode tiled > gpurun_out/r2_mode_tiled2.json 2> gpurun_out/r2_mode_tiled2.err; tail -c 300 gpurun_out/r2_mode_tiled2.err; head -c 1200 gpurun_out/r2_mode_tiled2.json
timeout 300 python tools/bench_widened.py --only superglue 2>&1 | tail -2 | cut -c1-300
DIMB_FUSE1A=1 timeout 240 ncu --set full --clock-control none --import-source on -k regex:conv1ab_pair_kernel -s 0 -c 1 -o gpurun_out/r2_prof_conv1ab8 -f python bench.py --quick --pairs 8 --steps 1 --warmup 3 > gpurun_out/ncu_fuse8.log 2>&1; tail -2 gpurun_out/ncu_fuse8.log
