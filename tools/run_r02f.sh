T=${1:-r02f}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print("value", d["value"], "e2e", d["e2e"])
print("aux", json.dumps(d["aux"])[:3000])
print("roofline", json.dumps(d["roofline"])[:1500])
PY
