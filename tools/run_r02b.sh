M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
python tools/view_probe.py --grid --out gpurun_out/r02b_views.json 2>&1 | tail -15
for v in c2 gridworst eye12km zoom5; do
  ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02b_launch_$v.csv python tools/view_probe.py --ncu $v --reps 2 > /dev/null 2>&1
done
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02b_launch_batch16.csv python tools/batch_sweep.py --once 16 > /dev/null 2>&1
ls -la gpurun_out | tail -8
