#!/bin/bash
# GPU call 5: SuperGlue on the tensor-core kernels, geometric verification, full suite with the new defaults, widened timings.
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "small|tiny|passed|failed|FAILED|rror|assert|max score delta" | cut -c1-300 | tee gpurun_out/r2_tests5.log | tail -30
timeout 300 python tools/bench_widened.py 2>&1 | tail -12 | cut -c1-400 | tee gpurun_out/r2_widened.log
timeout 300 python bench.py --quick --steps 10 --warmup 3 | tee gpurun_out/r2_quick5.json
