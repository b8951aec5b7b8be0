#!/bin/bash
# GPU call 3: regression of the elect.sync issue pattern (full suite), same-box A/B against the lane-0 build, bench.
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "pair [0-9]:|SuperPoint worst|tile [0-9]:|assembled|matches of 8192|passed|failed|FAILED|Error|rror:|differs" | cut -c1-300 > gpurun_out/r2_tests3.log; tail -40 gpurun_out/r2_tests3.log
for rep in 1 2; do
  DIMB_LIB=$PWD/deep-image-matching_b200/libdimb200_noelect.so timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2_ab_noelect$rep.json 2>/dev/null; cat gpurun_out/r2_ab_noelect$rep.json
  timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2_ab_elect$rep.json 2>/dev/null; cat gpurun_out/r2_ab_elect$rep.json
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; tail -c 300 gpurun_out/r2_bench3.err; python - <<'PY'
import json
j = json.load(open("gpurun_out/r2_bench3.json"))
print(j["value"], j["e2e"]["value"], j["clocks"])
for k, v in list(j["kernels"].items())[:14]:
    print(f"{k:20s} {v['ms_per_step']:8.3f} {v['tflops_algorithmic']}")
PY
