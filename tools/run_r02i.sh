T=${1:-r02i}
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; tail -2 gpurun_out/${T}_bench_ref.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "launches", d["gpu_launches"])
print("e2e", {k:v for k,v in d["e2e"].items() if not k.endswith("note")})
print("host", d["aux"]["host_enqueue_us_per_panorama"], d["aux"]["host_enqueue_fraction_of_device_period"])
print("c5", {k:v for k,v in d["aux"]["c5_grid"].items() if k!="what"})
print("special", d["aux"]["lone_ms_special_views"])
print("c3", d["aux"]["c3_sweep"])
print("cpu", d.get("cpu_baseline"))
print("roof", d["roofline"]["frac"], d["roofline"]["issue"]["frac"], d["roofline"]["latency_ms_single_panorama"])
r=json.load(open("gpurun_out/${T}_bench_ref.json"))
print("ref", r["value"], r["cpu_baseline"]["sample"][:200])
PY
