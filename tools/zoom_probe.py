import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import horizonator_b200 as hz
from tools import synth
C2_LAT, C2_LON = 34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0
tiles = synth.config2_tiles("/tmp/hz_tiles_c2")
h = hz.horizonator(C2_LAT, C2_LON, 3600, 600, SRTM1=True, dir_dems=tiles, render_radius_m=150000.)
for span in (10., 2., 0.5, 0.1, 0.02):
    for rep in range(2):
        t0 = time.perf_counter()
        img, rng = h.render(45. - span / 2, 45. + span / 2, znear=100., zfar=150000.)
        dt = time.perf_counter() - t0
    print("span %.2f deg: %.2f ms, terrain %.3f" % (span, dt * 1e3, (rng > 0).mean()), h.last_render_stats(), flush=True)
for span in (30., 10., 5.):
    h.pan_zoom(45. - span / 2, 45. + span / 2)
    h.profile(True); h.profile_read()
    for rep in range(3):
        h.render(45. - span / 2, 45. + span / 2, znear=100., zfar=150000.)
    p = h.profile_read(); c = h.render_counters(); h.profile(False)
    print("span", span, {k: round(v * 1e3, 1) for k, v in p.items() if k != "renders"}, c, flush=True)
