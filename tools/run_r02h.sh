T=${1:-r02h}
python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err; tail -5 gpurun_out/${T}_bench_2gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_2gpu.json"))
print("value", d["value"], "e2e", json.dumps(d["e2e"])[:900])
print("c5", json.dumps(d["aux"]["c5_grid"])[:700])
print("c4", json.dumps(d["aux"]["c4_wedge"]))
PY
