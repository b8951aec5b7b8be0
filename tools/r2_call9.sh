#!/bin/bash
# GPU call 9: TMEM-A probe, attention v5 (P in tensor memory) parity + A/B against v4, bit-mask NMS parity + A/B, fused conv1a as the default, full suite.
set -x
mkdir -p gpurun_out
timeout 120 python tools/probe_tmem_a.py 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k 'superpoint or nms or aliked' 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_nms2_tests.log
DIMB_ATTN=5 timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py tests/test_fast_mode.py -m gpu -q -x -k "lightglue or lg or pipe or superglue or cfg2 or chain or fast" 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_attn5_tests.log
for rep in 1 2; do
  for v in 4 5; do
    DIMB_ATTN=$v timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_ab_attn${v}_f1_$rep.json 2>gpurun_out/r2_ab_attn${v}_f1_$rep.err; cat gpurun_out/r2_ab_attn${v}_f1_$rep.json; tail -c 200 gpurun_out/r2_ab_attn${v}_f1_$rep.err
  done
done
for rep in 1 2; do DIMB_NMS=1 timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_ab_nms1_$rep.json 2>gpurun_out/r2_ab_nms1_$rep.err; cat gpurun_out/r2_ab_nms1_$rep.json; done
DIMB_ATTN=5 timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:lg_attn5_kernel -s 2 -c 1 -o gpurun_out/r2_prof_attn5 -f python bench.py --quick --pairs 8 --steps 1 --warmup 3 > gpurun_out/ncu_attn5.log 2>&1; tail -2 gpurun_out/ncu_attn5.log
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2_tests9.log
