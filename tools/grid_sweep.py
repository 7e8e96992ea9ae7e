"""Throughput over the 8x8 viewpoint grid (C5 flavour) for the current HORIZONATOR_* environment: batches of 16 in
flight, plus the median/maximum lone-render time."""
import os, sys, time
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
views = [(33.5 + (j + 0.5) / g + 1.0 / 7200.0, -117.5 + (i + 0.5) / g + 1.0 / 7200.0, -180.05, 179.95) for j in range(g) for i in range(g)]
B = 16
d_img = torch.empty((B, 600, 3600, 3), dtype=torch.uint8, device="cuda")
d_rng = torch.empty((B, 600, 3600), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(0, len(views), B):
        h.render_batch_device(views[k:k + B], d_img.data_ptr(), d_rng.data_ptr(), st)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
lone = []
for v in views[::3]:
    h.render_batch_device([v], d_img.data_ptr(), d_rng.data_ptr(), st); torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(3):
        h.render_batch_device([v], d_img.data_ptr(), d_rng.data_ptr(), st)
    torch.cuda.synchronize(); lone.append((time.perf_counter() - t1) / 3 * 1e3)
print("%-40s grid %.0f pano/s   lone median %.3f max %.3f ms" % (os.environ.get("HZ_TAG", ""), len(views) / dt, float(np.median(lone)), max(lone)))
