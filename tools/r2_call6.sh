#!/bin/bash
# GPU call 6: conv1a fused into the CTA-pair conv1b (DIMB_FUSE1A=1): parity, then same-box A/B.
set -x
mkdir -p gpurun_out
DIMB_FUSE1A=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "superpoint" 2>&1 | tail -12 | cut -c1-250 | tee gpurun_out/r2_fuse_tests.log
if grep -q "passed" gpurun_out/r2_fuse_tests.log && ! grep -q "failed\|rror" gpurun_out/r2_fuse_tests.log; then
  DIMB_FUSE1A=1 timeout 900 python -m pytest tests/test_cfg_parity.py tests/test_fast_mode.py -m gpu -q -s -k "cfg2_pipe or fast_mode" 2>&1 | grep -E "worst|passed|failed|rror" | cut -c1-250 | tee -a gpurun_out/r2_fuse_tests.log
  for rep in 1 2; do
    for f in 0 1; do
      DIMB_FUSE1A=$f timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2_ab_fuse${f}_$rep.json 2>gpurun_out/r2_ab_fuse${f}_$rep.err; cat gpurun_out/r2_ab_fuse${f}_$rep.json; tail -c 200 gpurun_out/r2_ab_fuse${f}_$rep.err
    done
  done
  DIMB_FUSE1A=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_fuse.json 2> gpurun_out/r2_bench_fuse.err; tail -c 300 gpurun_out/r2_bench_fuse.err
  python - <<'PY'
import json
j = json.load(open("gpurun_out/r2_bench_fuse.json"))
print(j["value"], j["e2e"]["value"], j["clocks"])
for k, v in list(j["kernels"].items())[:12]:
    print(f"{k:20s} {v['ms_per_step']:8.3f} {v['tflops_algorithmic']}")
PY
  DIMB_FUSE1A=1 timeout 240 ncu --set full --clock-control none --import-source on -k regex:conv1ab_pair_kernel -s 0 -c 1 -o gpurun_out/r2_prof_conv1ab -f python bench.py --quick --pairs 8 --steps 1 --warmup 3 > gpurun_out/ncu_fuse.log 2>&1; tail -2 gpurun_out/ncu_fuse.log
fi
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "superglue" 2>&1 | grep -E "small|tiny|passed|failed|rror" | cut -c1-200
timeout 300 python tools/bench_widened.py --only superglue 2>&1 | tail -3 | cut -c1-400
timeout 300 python bench.py --quick --steps 10 --warmup 3
