#!/usr/bin/env python
"""Generates tests/golden/llvmpipe_*.npz: renders of the REFERENCE ITSELF on a REAL OpenGL driver.

Run in the development container, where /root/reference exists:

    python tests/golden/make_golden_llvmpipe.py

oracle/_ref/libhorizonator_mesa.so is /root/reference/horizonator-lib.c + dem.c compiled UNMODIFIED and linked
to the Mesa 18.1.9 llvmpipe libGL that ships inside the image (with Nsight Compute), on the display-less Xlib of
oracle/mesa/fakex11.c and the GLX-pbuffer freeglut of oracle/mesa/glut_glx.c (recipe: oracle/Makefile, target
"mesa").  Shader compilation, clipping, rasterisation, the depth buffer and the read-back are Mesa's; nothing of
oracle/gl_pipeline.c is involved.  These files are what pins the oracle's GL rules F1-F9 (tests/test_llvmpipe.py)
and, on the GPU, the CUDA path itself (tests/test_gpu_parity.py).

The scenes of make_golden.py are repeated here on llvmpipe, plus larger ones that cover what small scenes cannot:
zoomed-in windows (triangles of hundreds of pixels), an eye high above the terrain, a wide full circle -- and the
two BASELINE workloads at FULL size (fullsize_c1_llvmpipe.npz: configs[0], 3600x300 over 2x2 SRTM3 tiles, 11.5 M
triangles, ~2 s on llvmpipe; fullsize_c2_llvmpipe.npz: configs[1], the benchmark panorama, 3600x600 over 150 km of
SRTM1, 274 M triangles, ~40 s on llvmpipe; ~0.6 MB each because most of a panorama is sky).
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import binding          # noqa: E402
from tools import synth             # noqa: E402

C1_LAT, C1_LON = 35.0 + 1.0 / 2400.0, -117.0 + 1.0 / 2400.0

SCENES = [
    # name,          W,   H,   R,   az0,     az1,    znear, zfar,    znc,  zfc,    lat,  lon,  viewer_z
    ("circle_small", 360, 60,  48,  -180.05, 179.95, 100., 100000., 100., 100000., None, None, None),
    ("quarter",      256, 96,  96,  30.0,    120.0,  50.,  20000.,  200., 10000.,  None, None, None),
    ("seam_odd_h",   300, 75,  64,  150.0,   210.0,  100., 40000.,  100., 40000.,  None, None, None),
    ("moved",        240, 80,  120, -60.0,   60.0,   100., 60000.,  100., 60000.,  C1_LAT - 0.02, C1_LON + 0.015, None),
    ("circle_wide",  900, 150, 400, -180.05, 179.95, 100., 100000., 100., 100000., None, None, None),
    ("zoom_5deg",    400, 200, 300, 41.0,    46.0,   100., 60000.,  100., 60000.,  None, None, None),
    ("high_eye",     480, 160, 300, -90.0,   90.0,   100., 80000.,  100., 80000.,  None, None, 3000.0),
]


def main():
    if not os.path.isdir("/root/reference"):
        raise SystemExit("needs /root/reference (run in the development container)")
    binding.build(ref=True)
    if not binding.have_mesa():
        raise SystemExit("no Mesa libGL in this image")
    tiles = synth.config1_tiles(os.path.join(tempfile.mkdtemp(prefix="hz_golden_"), "c1"))
    summary = {}
    for name, W, H, R, az0, az1, zn, zf, znc, zfc, lat, lon, vz in SCENES:
        r = binding.MesaReference(C1_LAT, C1_LON, W, H, dir_dems=tiles, render_radius_cells=R, viewer_z=vz, threads=4)
        kw = {} if lat is None else dict(lat=lat, lon=lon)
        img, rng = r.render(az0, az1, znear=zn, zfar=zf, znear_color=znc, zfar_color=zfc, **kw)
        version, renderer = r.gl_strings()
        assert "llvmpipe" in renderer, renderer
        assert (rng > 0).mean() > 0.01, name
        np.savez_compressed(os.path.join(HERE, "llvmpipe_%s.npz" % name), image=img, ranges=rng,
                            params=np.array([W, H, R, az0, az1, zn, zf, znc, zfc,
                                             -1000. if lat is None else lat, -1000. if lon is None else lon,
                                             -1. if vz is None else vz], np.float64),
                            viewer_z=np.float32(r.viewer_z))
        summary[name] = dict(viewer_z_at_init=float(r.viewer_z), hit_fraction=float((rng > 0).mean()))
        r.close()

    # horizonator_move()'s automatic eye height on llvmpipe: must equal what the fake-GL build produced (move.json)
    moves = json.load(open(os.path.join(HERE, "move.json")))
    r = binding.MesaReference(C1_LAT, C1_LON, 64, 16, dir_dems=tiles, render_radius_cells=600, threads=1)
    assert float(r.viewer_z) == moves[0]["viewer_z"]
    for m in moves[1:]:
        assert float(r.move(m["lat"], m["lon"])) == m["viewer_z"], m
    r.close()

    # horizonator_pick() (horizonator-lib.c:1216-1296: a one-pixel depth read-back + unproject) on the wide circle
    import ctypes as C
    PW, PH, PR = 900, 150, 400
    r = binding.MesaReference(C1_LAT, C1_LON, PW, PH, dir_dems=tiles, render_radius_cells=PR, threads=4)
    img, rng = r.render(-180.05, 179.95, znear=100., zfar=100000.)
    ys, xs = np.where(rng > 0)
    idx = np.random.default_rng(11).choice(len(ys), 60, replace=False)
    picks = []
    for x, y in [(int(xs[k]), int(ys[k])) for k in idx] + [(10, 5), (450, 0), (899, 149), (0, 149)]:
        la, lo = C.c_float(), C.c_float()
        ok = bool(r.lib().horizonator_pick(C.byref(r.ctx), C.byref(la), C.byref(lo), x, y))
        picks.append(dict(x=x, y=y, ok=ok, lat=float(la.value) if ok else None, lon=float(lo.value) if ok else None))
    r.close()
    assert sum(p["ok"] for p in picks) >= 60 and not all(p["ok"] for p in picks)

    # the two BASELINE workloads at full size
    C2_LAT, C2_LON = 34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0
    tiles2 = synth.config2_tiles(os.environ.get("HZ_BENCH_TILES", "/tmp/hz_tiles_c2"))
    for name, lat0, lon0, W, H, kw_init, zf, t in (
            ("c1", C1_LAT, C1_LON, 3600, 300, dict(dir_dems=tiles, render_radius_cells=1200), 100000., tiles),
            ("c2", C2_LAT, C2_LON, 3600, 600, dict(dir_dems=tiles2, render_radius_m=150000., SRTM1=True), 150000., tiles2)):
        r = binding.MesaReference(lat0, lon0, W, H, threads=os.cpu_count() or 1, **kw_init)
        img, rng = r.render(-180.05, 179.95, znear=100., zfar=zf)
        np.savez_compressed(os.path.join(HERE, "fullsize_%s_llvmpipe.npz" % name), image=img, ranges=rng,
                            viewer_z=np.float32(r.viewer_z))
        summary["fullsize_" + name] = dict(viewer_z_at_init=float(r.viewer_z), hit_fraction=float((rng > 0).mean()))
        r.close()

    with open(os.path.join(HERE, "llvmpipe.json"), "w") as f:
        json.dump(dict(gl_version=version, gl_renderer=renderer, move_json_reproduced=len(moves), scenes=summary,
                       pick=dict(scene=[PW, PH, PR, -180.05, 179.95, 100., 100000.], points=picks)),
                  f, indent=0, separators=(",", ":"))
        f.write("\n")
    print("llvmpipe golden renders written to", HERE, "--", version, "/", renderer)


if __name__ == "__main__":
    main()
