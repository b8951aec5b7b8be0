#!/bin/bash
# GPU call 1 of round 2: parity at the BASELINE sizes, bench with the reference legs, evidence of the shipped kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv | tail -1
timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -60 > gpurun_out/r2_tests1.log; tail -5 gpurun_out/r2_tests1.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -c 400 gpurun_out/r2_bench1.err; head -c 600 gpurun_out/r2_bench1.json
B="python bench.py --quick --pairs 8 --steps 1 --warmup 3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 465 -c 155 --csv --log-file gpurun_out/r2_launches_p8.csv $B > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
for spec in "conv1b:tc_gemm_pers_kernel:1" "attn:lg_attn3_kernel:1" "qk:EpiQK:1" "lngelu:lg_ln_gelu_kernel:1" "conv1a:sp_conv1a_kernel:1" "nms:sp_nms_kernel:1"; do
  IFS=: read name pat skip <<< "$spec"
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -o gpurun_out/r2_prof_$name -f $B > gpurun_out/ncu_$name.log 2>&1
  tail -2 gpurun_out/ncu_$name.log
done
ls -la gpurun_out | tail -12
