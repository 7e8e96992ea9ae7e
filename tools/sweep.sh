#!/bin/bash
# parameter sweep of the band structure / occlusion-box limits / lanes on the C2 benchmark (GPU box)
# each config: near_rings|bands|tile_pix|block_pix|batch
for cfg in "$@"; do
  IFS="|" read nr bands tp bp batch sp <<< "$cfg"
  HORIZONATOR_NEAR_RINGS=$nr HORIZONATOR_BANDS=$bands HORIZONATOR_OCCL_TILE_PIX=$tp HORIZONATOR_OCCL_BLOCK_PIX=$bp HORIZONATOR_LANES=$batch HORIZONATOR_SMALL_PIX=${sp:-64} \
    python bench.py --no-cpu-baseline --steps 40 --batch $batch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; k={a.split(' ')[0]: b for a, b in r['stage_ms_single_panorama'].items()}; c=d['aux']['culling']
print('$cfg', 'enq %.1f us' % d['aux']['host_enqueue_us_per_panorama'], 'value %.0f lat %.0f us e2e %.0f batch_e2e %.0f' % (d['value'], r['latency_ms_single_panorama']*1e3, d['e2e']['value'], d['e2e']['batch_call_value']), {a: round(b*1e3,1) for a,b in k.items()}, 'meshed', c['blocks_meshed'], 'blocks', c['blocks'], 'tris', c['triangles'])"
done
