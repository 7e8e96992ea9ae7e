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
for span in (360., 90., 30., 10., 5., 2.):
    v = [(C2_LAT, C2_LON, 45. - span / 2, 45. + span / 2)]
    for _ in range(2):
        h.render_batch_device(v, d_img.data_ptr(), d_rng.data_ptr(), st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        h.render_batch_device(v, d_img.data_ptr(), d_rng.data_ptr(), st)
    torch.cuda.synchronize()
    print("span %6.1f: %.3f ms (graphs=%s)" % (span, (time.perf_counter() - t0) / 5 * 1e3, os.environ.get("HORIZONATOR_GRAPHS", "1")), h.last_render_stats(), flush=True)
