#!/bin/bash
# GPU call 11: fused FFN0 + LayerNorm + GELU kernel (EpiFfnLn, gemm.cuh kFullRow) parity + A/B; 128 x 256 tiles for q/k + FFN0 as default.
set -x
mkdir -p gpurun_out
DIMB_FUSE_FFN=1 timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py tests/test_fast_mode.py -m gpu -q -x -k "lightglue or lg or pipe or cfg2 or chain or fast" 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/r2_fuseffn_tests.log
for rep in 1 2; do
  for f in 0 1; do
    DIMB_FUSE_FFN=$f timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_ab11_ffn${f}_$rep.json 2>gpurun_out/r2_ab11_ffn${f}_$rep.err; cut -c1-250 gpurun_out/r2_ab11_ffn${f}_$rep.json; tail -c 200 gpurun_out/r2_ab11_ffn${f}_$rep.err
  done
done
DIMB_FUSE_FFN=1 timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiFfnLn -s 2 -c 1 -o gpurun_out/r2_prof_ffnln -f python bench.py --quick --pairs 8 --steps 1 --warmup 2 > gpurun_out/ncu_ffnln.log 2>&1; tail -2 gpurun_out/ncu_ffnln.log
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiQK -s 2 -c 1 -o gpurun_out/r2_prof_qk256 -f python bench.py --quick --pairs 8 --steps 1 --warmup 2 > gpurun_out/ncu_qk256.log 2>&1; tail -2 gpurun_out/ncu_qk256.log
