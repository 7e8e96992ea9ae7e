probe() { python tools/view_probe.py c2 gridmedian gridworst 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    n, _, j = l.partition(' ')
    try: d = json.loads(j)
    except Exception: continue
    print('  %-10s %.4f ms' % (n, d['lone_ms']), d['stage_us'])"; }
echo "== fetch16 (default)"; probe
echo "== fetch8"; HORIZONATOR_LIBRARY=$PWD/horizonator_b200/lib/libhorizonator_fetch8.so probe
echo "== fetch32"; HORIZONATOR_LIBRARY=$PWD/horizonator_b200/lib/libhorizonator_fetch32.so probe
for g in 50 100 250; do echo "== GRID_SCALE=$g"; HORIZONATOR_GRID_SCALE=$g probe; done
for b in "32" "64" "24,72" "16,48,110"; do echo "== BANDS=$b"; HORIZONATOR_BANDS=$b HORIZONATOR_BANDS_BATCH=10,24,56,120 probe; done
echo "== batch with fetch variants"
python tools/batch_sweep.py --reps 6 --batches 64 "" | tail -1
HORIZONATOR_LIBRARY=$PWD/horizonator_b200/lib/libhorizonator_fetch8.so python tools/batch_sweep.py --reps 6 --batches 64 "" | tail -1
HORIZONATOR_LIBRARY=$PWD/horizonator_b200/lib/libhorizonator_fetch32.so python tools/batch_sweep.py --reps 6 --batches 64 "" | tail -1
