T=${1:-r02k}
python tools/batch_sweep.py --reps 6 --batches 64 --out gpurun_out/${T}_sweep2.jsonl "BANDS_BATCH=10:24:56:120,OCCL_TILE_PIX=256,OCCL_BLOCK_PIX=64" "BANDS_BATCH=10:24:56:120,OCCL_TILE_PIX=512,OCCL_BLOCK_PIX=128" "BANDS_BATCH=8:20:48:110,OCCL_TILE_PIX=256,OCCL_BLOCK_PIX=64" "BANDS_BATCH=10:24:56:120:200,OCCL_TILE_PIX=256,OCCL_BLOCK_PIX=64" "OCCL_TILE_PIX=512,OCCL_BLOCK_PIX=128" "OCCL_TILE_PIX=1024,OCCL_BLOCK_PIX=64" "BANDS_BATCH=10:24:56:120,OCCL_TILE_PIX=256,OCCL_BLOCK_PIX=64,SETS=3" 2>&1 | tail -8
for cfg in "" "HORIZONATOR_OCCL_TILE_PIX=256 HORIZONATOR_OCCL_BLOCK_PIX=64" "HORIZONATOR_OCCL_TILE_PIX=256 HORIZONATOR_OCCL_BLOCK_PIX=64 HORIZONATOR_BANDS=24,72" "HORIZONATOR_OCCL_TILE_PIX=256 HORIZONATOR_OCCL_BLOCK_PIX=64 HORIZONATOR_BANDS=10,24,56,120"; do
  echo "== lone views with: $cfg"
  env $cfg python tools/view_probe.py c2 gridworst gridmedian eye12km zoom10 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    n, _, j = l.partition(' ')
    try: d = json.loads(j)
    except Exception: continue
    print('  %-10s %.4f ms' % (n, d['lone_ms']), d['stage_us'])"
done
