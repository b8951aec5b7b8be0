python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
python bench.py --precision fast --no-cpu-baseline > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err
python - <<PY
import json
for f in ("bench_exact","bench_fast"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["value"], d["e2e"]["value"], d["e2e"]["u8_images"]["value"], d["ms_per_step"], d["clocks"], d["adaptive"]["value"], d["roofline"]["achieved"], d["roofline_whole_step"]["achieved"], d.get("cpu_baseline",{}).get("value"))
PY
