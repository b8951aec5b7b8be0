#!/bin/bash
# GPU call 14: fused conv1a with 12 producer warps (two concurrent channel-half groups of 6 x 30 lanes): SuperPoint parity + kernel time.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py -m gpu -q -x -k "superpoint or pipe or cfg2 or chain" 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/r2_prod12_tests.log
for rep in 1 2; do
  timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_q14_$rep.json 2>gpurun_out/r2_q14_$rep.err; cut -c1-900 gpurun_out/r2_q14_$rep.json; tail -c 200 gpurun_out/r2_q14_$rep.err
  DIMB_FUSE1A=0 timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_q14_nofuse_$rep.json 2>gpurun_out/r2_q14_nofuse_$rep.err; cut -c1-900 gpurun_out/r2_q14_nofuse_$rep.json
done
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:conv1ab_pair_kernel -s 0 -c 1 -o gpurun_out/r2_prof_conv1ab12 -f python bench.py --quick --pairs 8 --steps 1 --warmup 2 > gpurun_out/ncu_conv1ab12.log 2>&1; tail -2 gpurun_out/ncu_conv1ab12.log
