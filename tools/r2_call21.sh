#!/bin/bash
# GPU call 21: conv1a on the tensor cores inside the fused CTA-pair kernel (conv1ab_mma_pair_kernel, DIMB_FUSE1A=2): parity + A/B + ncu.
set -x
mkdir -p gpurun_out
DIMB_FUSE1A=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_cfg_parity.py -m gpu -q -x -k "superpoint or pipe or cfg2 or chain" 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/r2_c1amma_tests.log
for rep in 1 2; do
  for f in 1 2; do
    DIMB_FUSE1A=$f timeout 300 python bench.py --quick --kernels --steps 10 --warmup 3 > gpurun_out/r2_q21_fuse${f}_$rep.json 2>gpurun_out/r2_q21_fuse${f}_$rep.err; python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2_q21_fuse${f}_$rep.json')); k=d['kernels_ms_per_step']
    print('fuse=$f', round(d['value'],1), 'conv1ab', k.get('sp.conv1ab'))
except Exception as e:
    print('fuse=$f failed', e); print(open('gpurun_out/r2_q21_fuse${f}_$rep.err').read()[-500:])
P
  done
done
DIMB_FUSE1A=2 timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:conv1ab_mma_pair_kernel -s 0 -c 1 -o gpurun_out/r2_prof_conv1ab_mma -f python bench.py --quick --pairs 8 --steps 1 --warmup 2 > gpurun_out/ncu_conv1ab_mma.log 2>&1; tail -2 gpurun_out/ncu_conv1ab_mma.log
