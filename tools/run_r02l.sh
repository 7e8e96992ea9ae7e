T=${1:-r02l}
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/batch_sweep.py --reps 6 --batches 16,64,256 --out gpurun_out/${T}_sweep.jsonl "" 2>&1 | tail -2
python tools/view_probe.py --out gpurun_out/${T}_views.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    n, _, j = l.partition(' ')
    try: d = json.loads(j)
    except Exception: continue
    print('  %-10s %.4f ms' % (n, d['lone_ms']), d['stage_us'])"
