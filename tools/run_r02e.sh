M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
T=${1:-r02e}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/view_probe.py --out gpurun_out/${T}_views.json 2>&1 | cut -c1-330
echo "--- mesh3 variant"
HORIZONATOR_LIBRARY=$PWD/horizonator_b200/lib/libhorizonator_mesh3.so python tools/view_probe.py c2 eye12km gridworst 2>&1 | cut -c1-330
for v in c2 gridworst eye12km zoom5 zoom10; do
  ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_launch_$v.csv python tools/view_probe.py --ncu $v --reps 2 > /dev/null 2>&1
done
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_launch_batch16.csv python tools/batch_sweep.py --once 16 > /dev/null 2>&1
python tools/batch_sweep.py --reps 20 --out gpurun_out/${T}_sweep.jsonl "" "SETS=4" "MID_PIX=0" "MID_PIX=128" 2>&1 | tail -4
HORIZONATOR_LIBRARY=$PWD/horizonator_b200/lib/libhorizonator_mesh3.so python tools/batch_sweep.py --reps 20 "" 2>&1 | tail -1
