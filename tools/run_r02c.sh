M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
T=${1:-r02c}
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python tools/view_probe.py --out gpurun_out/${T}_views.json c2 eye12km zoom5 gridworst gridmedian 2>&1 | cut -c1-400
for v in c2 gridworst eye12km zoom5; do
  ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_launch_$v.csv python tools/view_probe.py --ncu $v --reps 2 > /dev/null 2>&1
done
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_launch_batch16.csv python tools/batch_sweep.py --once 16 > /dev/null 2>&1
python tools/batch_sweep.py --reps 20 --out gpurun_out/${T}_sweep.jsonl "" "SETS=4" 2>&1 | tail -3
