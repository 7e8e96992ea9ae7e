"""One mid-size render through every kernel (near pass + two bands + large triangles + lanes), for compute-sanitizer."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import horizonator_b200 as hz
from tools import synth
d = synth.config1_tiles(os.path.join(tempfile.gettempdir(), "hz_smoke_tiles"))
lat, lon = 35.0 + 1.0 / 2400.0, -117.0 + 1.0 / 2400.0
os.environ.setdefault("HORIZONATOR_BANDS", "6")
h = hz.horizonator(lat, lon, 720, 120, dir_dems=d, render_radius_cells=400)
a = h.render(-180.05, 179.95, zfar=100000.)
h.set_zextents(100., 100000.)
bi, br = h.render_batch([(lat, lon, -180.05, 179.95), (lat + 0.01, lon, -90., 90.), (lat, lon + 0.01, 0., 45.)])
assert np.array_equal(bi[0], a[0]) and np.array_equal(br[0], a[1])
print("ok", h.last_render_stats())
