#!/bin/bash
# GPU call 10: attention v6 (v5 + two softmax threads per row) parity + A/B, 128 x 256 LightGlue GEMM tiles parity + A/B,
# full bench on the new defaults (attention v5, bit-mask NMS, fused conv1a), ncu of the new kernels + launch list.
set -x
mkdir -p gpurun_out
timeout 120 python tools/probe_tmem_a.py 2>&1 | tail -2
DIMB_ATTN=6 timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py tests/test_fast_mode.py -m gpu -q -x -k "lightglue or lg or pipe or superglue or cfg2 or chain or fast" 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_attn6_tests.log
DIMB_BN256=1 timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py tests/test_fast_mode.py -m gpu -q -x -k "lightglue or lg or pipe or cfg2 or chain or fast" 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_bn256_tests.log
for rep in 1 2; do
  for v in 5 6; do
    DIMB_ATTN=$v timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_ab10_attn${v}_$rep.json 2>gpurun_out/r2_ab10_attn${v}_$rep.err; cut -c1-250 gpurun_out/r2_ab10_attn${v}_$rep.json; tail -c 200 gpurun_out/r2_ab10_attn${v}_$rep.err
  done
  DIMB_BN256=1 timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_ab10_bn256_$rep.json 2>gpurun_out/r2_ab10_bn256_$rep.err; cut -c1-250 gpurun_out/r2_ab10_bn256_$rep.json; tail -c 200 gpurun_out/r2_ab10_bn256_$rep.err
done
timeout 600 python bench.py > gpurun_out/r2_bench_call10.json 2> gpurun_out/r2_bench_call10.err; tail -c 300 gpurun_out/r2_bench_call10.err; head -c 400 gpurun_out/r2_bench_call10.json
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:sp_nms2_kernel -s 1 -c 1 -o gpurun_out/r2_prof_nms2 -f python bench.py --quick --pairs 8 --steps 1 --warmup 2 > gpurun_out/ncu_nms2.log 2>&1; tail -2 gpurun_out/ncu_nms2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_call10.csv python bench.py --quick --pairs 8 --steps 2 --warmup 1 > gpurun_out/ncu_launches10.log 2>&1; tail -1 gpurun_out/ncu_launches10.log
