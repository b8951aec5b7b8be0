#!/bin/bash
# Final evidence of round 2 (one gpurun call): full GPU suite, the bench line, widened bench, launch list, compute-sanitizer passes.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2_tests_final.log
timeout 700 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 300 gpurun_out/r2_bench_final.err; head -c 1500 gpurun_out/r2_bench_final.json
timeout 400 python tools/bench_widened.py > gpurun_out/r2_widened_final.jsonl 2> gpurun_out/r2_widened_final.err; tail -c 200 gpurun_out/r2_widened_final.err; cut -c1-260 gpurun_out/r2_widened_final.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --quick --pairs 8 --steps 2 --warmup 1 > gpurun_out/ncu_launches_final.log 2>&1; tail -1 gpurun_out/ncu_launches_final.log
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/r2_compute_sanitizer_smoke.log 2>&1; tail -4 gpurun_out/r2_compute_sanitizer_smoke.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lighterglue_trained or nms_is_exact or tiny or superglue_matches" > gpurun_out/r2_compute_sanitizer_tests.log 2>&1; tail -5 gpurun_out/r2_compute_sanitizer_tests.log
timeout 600 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/r2_compute_sanitizer_racecheck_smoke.log 2>&1; tail -4 gpurun_out/r2_compute_sanitizer_racecheck_smoke.log
