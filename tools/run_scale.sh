# bench.py on N GPUs of one box (torchrun), for each N given: results into gpurun_out/<tag>_bench_<N>gpu.json
T=$1; shift
for N in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
  tail -2 gpurun_out/${T}_bench_${N}gpu.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_${N}gpu.json"))
print("N=$N value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pageable", round(d["e2e"]["pageable_host_buffers_value"]), "batch", round(d["e2e"]["batch_call_value"]), "ceiling", round(d["e2e"]["d2h_ceiling_value"]))
print("   c5", round(d["aux"]["c5_grid"]["value"]), "per gpu", d["aux"]["c5_grid"]["viewpoints_per_gpu"])
print("   c4", json.dumps(d["aux"]["c4_wedge"]))
PY
done
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1; head -14 gpurun_out/${T}_topo.txt
