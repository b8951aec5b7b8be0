#!/bin/bash
# GPU call 12: epilogue-preload build: parity (LightGlue / SuperGlue / NN / ALIKED users of the transposing epilogues), quick bench with
# kernel table, the secondary workloads (exhaustive = cfg4 shape, nn = cfg5, tiled = cfg3), widened bench (LighterGlue / SuperGlue).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2_tests12.log
for rep in 1 2; do
  timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_q12_$rep.json 2>gpurun_out/r2_q12_$rep.err; cut -c1-2000 gpurun_out/r2_q12_$rep.json; tail -c 200 gpurun_out/r2_q12_$rep.err
done
timeout 400 python bench.py --mode exhaustive > gpurun_out/r2_mode_exhaustive.json 2> gpurun_out/r2_mode_exhaustive.err; tail -c 300 gpurun_out/r2_mode_exhaustive.err; head -c 1500 gpurun_out/r2_mode_exhaustive.json
timeout 300 python bench.py --mode nn > gpurun_out/r2_mode_nn.json 2> gpurun_out/r2_mode_nn.err; tail -c 300 gpurun_out/r2_mode_nn.err; head -c 1500 gpurun_out/r2_mode_nn.json
timeout 300 python bench.py --mode tiled > gpurun_out/r2_mode_tiled.json 2> gpurun_out/r2_mode_tiled.err; tail -c 300 gpurun_out/r2_mode_tiled.err; head -c 2000 gpurun_out/r2_mode_tiled.json
timeout 400 python tools/bench_widened.py > gpurun_out/r2_widened.json 2> gpurun_out/r2_widened.err; tail -c 300 gpurun_out/r2_widened.err; head -c 3000 gpurun_out/r2_widened.json
