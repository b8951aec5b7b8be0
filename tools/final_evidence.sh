#!/bin/bash
# The one GPU call that regenerates the evidence under profiles/ (copy from gpurun_out/ afterwards, see profiles/README.md).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err; tail -c 300 gpurun_out/bench_exact.err
python bench.py --precision fast --no-cpu-baseline > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
B="python bench.py --quick --pairs 8 --steps 1 --warmup 3"
# one fixed-work step at 8 pairs = 155 launches; 4 steps run (3 warm-up + 1): skip 465, capture the last step
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 465 -c 155 --csv --log-file gpurun_out/launches_p8.csv $B > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
if [ "$1" = "full" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_pers_kernel -c 1 -o gpurun_out/prof_conv1b -f $B > gpurun_out/ncu_c1b.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sp_conv1a_kernel|sp_nms_kernel" -c 2 -o gpurun_out/prof_c1a_nms2 -f $B > gpurun_out/ncu_c1a.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:lg_attn3_kernel -c 1 -o gpurun_out/prof_attn -f $B > gpurun_out/ncu_attn.log 2>&1
fi
python tools/bench_widened.py 2>&1 | tail -7 | cut -c1-300
ls -la gpurun_out | tail -15
