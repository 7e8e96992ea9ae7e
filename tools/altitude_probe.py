"""Device time of a full-circle C2 panorama with the eye at given heights above sea level (C API only: explicit viewer_z)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import horizonator_b200 as hz
from tools import synth
C2_LAT, C2_LON = 34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0
tiles = synth.config2_tiles("/tmp/hz_tiles_c2")
h = hz.horizonator(C2_LAT, C2_LON, 3600, 600, SRTM1=True, dir_dems=tiles, render_radius_m=150000.)
h.set_zextents(100., 150000.)
d_img = torch.empty((600, 3600, 3), dtype=torch.uint8, device="cuda"); d_rng = torch.empty((600, 3600), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for z in (-1., 3000., 6000., 12000.):
    v = [(C2_LAT, C2_LON, -180.05, 179.95, z)]
    for _ in range(2):
        h.render_batch_device(v, d_img.data_ptr(), d_rng.data_ptr(), st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        h.render_batch_device(v, d_img.data_ptr(), d_rng.data_ptr(), st)
    torch.cuda.synchronize()
    print("eye at %7.0f m: %.3f ms, terrain %.3f" % (z, (time.perf_counter() - t0) / 5 * 1e3, float((d_rng > 0).float().mean())), h.last_render_stats(), flush=True)
