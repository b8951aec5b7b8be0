#!/bin/bash
set -x
mkdir -p gpurun_out
for rep in 1 2; do
  for pair in 0 1; do
    DIMB_PAIR=$pair timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2_ab_pair${pair}_$rep.json 2>gpurun_out/r2_ab_pair${pair}_$rep.err; cat gpurun_out/r2_ab_pair${pair}_$rep.json; tail -c 200 gpurun_out/r2_ab_pair${pair}_$rep.err
  done
done
DIMB_PAIR=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_pair.json 2> gpurun_out/r2_bench_pair.err; tail -c 300 gpurun_out/r2_bench_pair.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2_bench_pair.json"))
print(j["value"], j["e2e"]["value"], j["clocks"])
for k, v in list(j["kernels"].items())[:16]:
    print(f"{k:20s} {v['ms_per_step']:8.3f} {v['tflops_algorithmic']}")
PY
DIMB_PAIR=1 timeout 240 ncu --set full --clock-control none --import-source on -k regex:conv64_pair_kernel -s 0 -c 1 -o gpurun_out/r2_prof_conv1b_pair -f python bench.py --quick --pairs 8 --steps 1 --warmup 3 > gpurun_out/ncu_pair.log 2>&1; tail -2 gpurun_out/ncu_pair.log
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiQK -s 1 -c 1 -o gpurun_out/r2_prof_qk_resb -f python bench.py --quick --pairs 8 --steps 1 --warmup 3 > gpurun_out/ncu_qk2.log 2>&1; tail -2 gpurun_out/ncu_qk2.log
DIMB_PAIR=1 timeout 1200 python -m pytest tests/test_cfg_parity.py tests/test_gpu_parity.py tests/test_store_and_sets.py -m gpu -q -s 2>&1 | grep -E "pair [0-9]:|worst|passed|failed|FAILED|rror|assert" | cut -c1-400 | tee gpurun_out/r2_tests4.log
