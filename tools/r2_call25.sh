#!/bin/bash
# GPU call 25: confirmation on the final build: full GPU suite, the bench line, widened bench.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -8 | cut -c1-400 | tee gpurun_out/r2_tests_final3.log
timeout 700 python bench.py > gpurun_out/r2_bench_final3.json 2> gpurun_out/r2_bench_final3.err; tail -c 300 gpurun_out/r2_bench_final3.err; head -c 700 gpurun_out/r2_bench_final3.json
timeout 400 python tools/bench_widened.py > gpurun_out/r2_widened_final3.jsonl 2> gpurun_out/r2_widened_final3.err; cut -c1-200 gpurun_out/r2_widened_final3.jsonl
