#!/bin/bash
# GPU call 2 of round 2: full parity suite, attention A/B, secondary bench modes, the missing ncu captures.
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2_tests2.log; tail -25 gpurun_out/r2_tests2.log | cut -c1-250
for lazy in 0 8; do
  DIMB_ATTN_LAZY=$lazy timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2_ab_lazy$lazy.json 2> gpurun_out/r2_ab_lazy$lazy.err; cat gpurun_out/r2_ab_lazy$lazy.json
done
timeout 300 python bench.py --mode nn > gpurun_out/r2_mode_nn.json 2> gpurun_out/r2_mode_nn.err; tail -c 300 gpurun_out/r2_mode_nn.err; head -c 1500 gpurun_out/r2_mode_nn.json
timeout 300 python bench.py --mode exhaustive > gpurun_out/r2_mode_exh.json 2> gpurun_out/r2_mode_exh.err; tail -c 300 gpurun_out/r2_mode_exh.err; head -c 1200 gpurun_out/r2_mode_exh.json
timeout 300 python bench.py --mode tiled > gpurun_out/r2_mode_tiled.json 2> gpurun_out/r2_mode_tiled.err; tail -c 300 gpurun_out/r2_mode_tiled.err; head -c 1500 gpurun_out/r2_mode_tiled.json
B="python bench.py --quick --pairs 8 --steps 1 --warmup 3"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_pers_kernel -s 0 -c 1 -o gpurun_out/r2_prof_conv1b -f $B > gpurun_out/ncu_conv1b.log 2>&1; tail -2 gpurun_out/ncu_conv1b.log
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiQK -s 1 -c 1 -o gpurun_out/r2_prof_qk -f $B > gpurun_out/ncu_qk.log 2>&1; tail -2 gpurun_out/ncu_qk.log
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiLgResidual -s 1 -c 1 -o gpurun_out/r2_prof_ffn3 -f $B > gpurun_out/ncu_ffn3.log 2>&1; tail -2 gpurun_out/ncu_ffn3.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:lg_attn3_kernel -s 1 -c 1 -o gpurun_out/r2_prof_attn_lazy -f $B > gpurun_out/ncu_attn2.log 2>&1; tail -2 gpurun_out/ncu_attn2.log
ls -la gpurun_out | tail -8
