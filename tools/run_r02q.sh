python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/batch_sweep.py --reps 10 --batches 64 "" | tail -1
python tools/view_probe.py c2 gridworst eye12km zoom10 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    n, _, j = l.partition(' ')
    try: d = json.loads(j)
    except Exception: continue
    print('  %-10s %.4f ms' % (n, d['lone_ms']), d['stage_us'])"
