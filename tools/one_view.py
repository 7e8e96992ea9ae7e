"""Renders one view of the C2 context a few times (for ncu):  python tools/one_view.py SPAN_DEG [AZ_CENTER]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import horizonator_b200 as hz
from tools import synth
C2_LAT, C2_LON = 34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0
span = float(sys.argv[1]); c = float(sys.argv[2]) if len(sys.argv) > 2 else 45.
tiles = synth.config2_tiles("/tmp/hz_tiles_c2")
h = hz.horizonator(C2_LAT, C2_LON, 3600, 600, SRTM1=True, dir_dems=tiles, render_radius_m=150000.)
h.set_zextents(100., 150000.)
d_img = torch.empty((600, 3600, 3), dtype=torch.uint8, device="cuda"); d_rng = torch.empty((600, 3600), dtype=torch.float32, device="cuda")
for _ in range(3):
    h.render_batch_device([(C2_LAT, C2_LON, c - span / 2, c + span / 2)], d_img.data_ptr(), d_rng.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
