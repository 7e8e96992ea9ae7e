"""Per-stage cost of single renders over the 8x8 viewpoint grid of tools/scenarios.py (C5 flavour): finds what the
expensive viewpoints spend their time on."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import horizonator_b200 as hz
from tools import synth
C2_LAT, C2_LON = 34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0
tiles = synth.config2_tiles(os.environ.get("HZ_BENCH_TILES", "/tmp/hz_tiles_c2"))
h = hz.horizonator(C2_LAT, C2_LON, 3600, 600, SRTM1=True, dir_dems=tiles, render_radius_m=150000.)
h.set_zextents(100., 150000.)
g = 8
views = [(C2_LAT, C2_LON, -180.05, 179.95)] + [(33.5 + (j + 0.5) / g + 1.0 / 7200.0, -117.5 + (i + 0.5) / g + 1.0 / 7200.0, -180.05, 179.95) for j in range(g) for i in range(g)]
d_img = torch.empty((600, 3600, 3), dtype=torch.uint8, device="cuda")
d_rng = torch.empty((600, 3600), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
rows = []
for v in views:
    for _ in range(2):
        h.render_batch_device([v], d_img.data_ptr(), d_rng.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(5):
        h.render_batch_device([v], d_img.data_ptr(), d_rng.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    h.profile(True); h.profile_read()
    for _ in range(3):
        h.render_batch_device([v], d_img.data_ptr(), d_rng.data_ptr(), st)
    torch.cuda.synchronize()
    p = h.profile_read(); c = h.render_counters(); s = h.last_render_stats(); h.profile(False)
    hit = float((d_rng > 0).float().mean().item())
    rows.append(dict(lat=v[0], lon=v[1], ms=ms, hit=hit, stages={k: round(p[k] * 1e3, 1) for k in ("prepare", "near", "big_near", "march", "big_far", "resolve")},
                     meshed=c["blocks_meshed"], blocks=c["blocks"], tris=c["triangles"], big=s["big_entries"], tiles_alive=c["tiles"] - c["tiles_far"] - c["tiles_window"] - c["tiles_occluded"]))
rows_sorted = sorted(rows[1:], key=lambda r: -r["ms"])
print("benchmark view:", json.dumps(rows[0]))
for r in rows_sorted[:5] + rows_sorted[-3:]:
    print(json.dumps(r))
print("median ms", float(np.median([r["ms"] for r in rows[1:]])))
