#!/usr/bin/env python
"""Generates the golden vectors of tests/golden/ from the REFERENCE ITSELF.

Run in the development container, where /root/reference exists:

    python tests/golden/make_golden.py

The reference ships no tests and no golden vectors, so these are outputs of the
reference's own code run here: oracle/_ref/libhorizonator_ref.so is /root/reference/horizonator-lib.c +
dem.c compiled UNMODIFIED (oracle/Makefile).  Everything the host side of the reference computes is
therefore pinned by the reference's own arithmetic:

  dem_geometry.json   horizonator_dem_init(): radius, origin tile/cell, tile counts, success/failure
  dem_samples.json    horizonator_dem_sample() on the seeded synthetic tiles (incl. voids, missing tile,
                      zero-length tile) and horizonator_dem_bounds_latlon_deg()
  move.json           horizonator_move(): automatic viewer height (via the read-back of a flat render)
  geometry.json       horizonator_x_from_az / _project / _unproject
  render_*.npz        horizonator_render_offscreen(): BGR image + range image of small scenes.  The GL
                      driver below the reference is the software restatement oracle/gl_pipeline.c: the host
                      logic, read-back, flips and the depth->range conversion are the reference's, the
                      rasterisation rules are the restatement's.  (The same scenes rendered by the reference
                      on a REAL driver, Mesa llvmpipe, are made by make_golden_llvmpipe.py.)

The synthetic tiles come from tools/synth_hgt.c (integer hash noise, seeded); tiles_sha256.json records
what the generator produced here, and the tests check that first.
"""
import ctypes as C
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
from horizonator_b200 import dem_context_t   # noqa: E402  (struct layout only)

C1_LAT, C1_LON = 35.0 + 1.0 / 2400.0, -117.0 + 1.0 / 2400.0

# (lat, lon, radius_cells, radius_m, SRTM1)
DEM_CASES = [
    (C1_LAT, C1_LON, 1200, -1.0, False),          # BASELINE config 1
    (C1_LAT, C1_LON, 1201, -1.0, False),          # origin cell (0,0): the reference's out-of-bounds corner
    (C1_LAT, C1_LON, 1, -1.0, False),
    (C1_LAT, C1_LON, 48, -1.0, False),
    (34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0, -1, 150000.0, True),    # BASELINE config 2
    (34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0, 7200, -1.0, True),      # SRTM1 maximum
    (34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0, 7201, -1.0, True),      # too large: fails
    (2.0 + 1.0 / 7200.0, 20.0 + 1.0 / 7200.0, -1, 150000.0, True),       # low latitude, eastern hemisphere
    (-33.3, 18.7, 500, -1.0, False),                                     # southern hemisphere
    (-0.2, -0.3, 700, -1.0, False),                                      # straddles equator and Greenwich
    (35.0, -117.0, 600, -1.0, False),                                    # viewer exactly on a tile corner
    (34.5, -117.5, -1, 40000.0, False),
    (34.5, -117.5, -1, -1.0, False),                                     # both radii < 0: fails
    (34.5, -117.5, 100, 5000.0, False),                                  # both radii > 0: fails
    (60.2, 10.4, -1, 30000.0, False),                                    # high latitude: narrow cells
]

RENDER_SCENES = [
    # name,          W,   H,  R,  az0,     az1,    znear, zfar,    znc,  zfc,   lat,  lon
    ("circle_small", 360, 60, 48, -180.05, 179.95, 100., 100000., 100., 100000., None, None),
    ("quarter",      256, 96, 96,   30.0,  120.0,  50.,  20000.,  200., 10000.,  None, None),
    ("seam_odd_h",   300, 75, 64,  150.0,  210.0,  100., 40000.,  100., 40000.,  None, None),
    ("moved",        240, 80, 120, -60.0,   60.0,  100., 60000.,  100., 60000.,  C1_LAT - 0.02, C1_LON + 0.015),
]


def dem_geometry(tmp):
    L = binding.Reference.lib()
    out = []
    for lat, lon, rc, rm, srtm1 in DEM_CASES:
        d = os.path.join(tmp, "empty_%d" % len(out))
        os.makedirs(d, exist_ok=True)
        # zero-length stand-in tiles: dem.c:210-222 accepts them silently as sea, so no 26 MB files are needed
        for la in range(int(np.floor(lat)) - 3, int(np.floor(lat)) + 4):
            for lo in range(int(np.floor(lon)) - 3, int(np.floor(lon)) + 4):
                open(os.path.join(d, synth.tile_name(la, lo)), "wb").close()
        ctx = dem_context_t()
        ok = bool(L.horizonator_dem_init(C.byref(ctx), lat, lon, rc, rm, os.fsencode(d), srtm1))
        rec = dict(lat=lat, lon=lon, radius_cells=rc, radius_m=rm, SRTM1=srtm1, ok=ok)
        if ok:
            rec.update(origin_dem_lon_lat=list(ctx.origin_dem_lon_lat), origin_dem_cellij=list(ctx.origin_dem_cellij),
                       Ndems_ij=list(ctx.Ndems_ij), R=ctx.radius_cells, cells_per_deg=ctx.cells_per_deg)
            b = [C.c_float() for _ in range(4)]
            L.horizonator_dem_bounds_latlon_deg.argtypes = [C.c_void_p] + [C.POINTER(C.c_float)] * 4
            L.horizonator_dem_bounds_latlon_deg(C.byref(ctx), *[C.byref(x) for x in b])
            rec["bounds_lat0_lon0_lat1_lon1"] = [float(np.float32(x.value)) for x in b]
            L.horizonator_dem_deinit(C.byref(ctx))
        out.append(rec)
    return out


def dem_samples(tiles, tiles_holes):
    L = binding.Reference.lib()
    out = []
    rs = np.random.default_rng(11)
    for name, d, R in (("c1", tiles, 1200), ("c1_small", tiles, 48), ("holes", tiles_holes, 900)):
        ctx = dem_context_t()
        assert L.horizonator_dem_init(C.byref(ctx), C1_LAT, C1_LON, R, -1.0, os.fsencode(d), False)
        N = 2 * R
        pts = [(0, 0), (N - 1, N - 1), (0, N - 1), (N - 1, 0), (-1, 5), (5, -1)]
        # both sides of every tile boundary of the mosaic (shared-edge rule, dem.c:287-291)
        oi, oj = ctx.origin_dem_cellij[0], ctx.origin_dem_cellij[1]
        for k in (1199, 1200, 1201):
            if 0 <= k - oi < N:
                pts += [(k - oi, int(rs.integers(0, N))) for _ in range(6)]
            if 0 <= k - oj < N:
                pts += [(int(rs.integers(0, N)), k - oj) for _ in range(6)]
        pts += [(int(rs.integers(0, N)), int(rs.integers(0, N))) for _ in range(400)]
        vals = [int(L.horizonator_dem_sample(C.byref(ctx), i, j)) for i, j in pts]
        out.append(dict(name=name, R=R, points=pts, values=vals))
        L.horizonator_dem_deinit(C.byref(ctx))
    return out


def geometry_vectors():
    L = binding.Reference.lib()
    rs = np.random.default_rng(5)
    d = C.c_double
    out = dict(x_from_az=[], project=[], unproject=[])
    for _ in range(60):
        az0 = float(rs.uniform(-400, 400)); span = float(rs.uniform(5, 359)); az = float(rs.uniform(-720, 720))
        W = int(rs.integers(16, 4000))
        x, per = d(), d()
        ok = bool(L.horizonator_x_from_az(C.byref(x), C.byref(per), np.radians(az), np.radians(az0), np.radians(az0 + span), W))
        out["x_from_az"].append(dict(az_rad=np.radians(az), az_rad0=np.radians(az0), az_rad1=np.radians(az0 + span),
                                     width=W, ok=ok, x=x.value if ok else None, per=per.value if ok else None))
    for _ in range(60):
        latv, lonv = float(rs.uniform(-60, 60)), float(rs.uniform(-179, 179))
        lat, lon = latv + float(rs.uniform(-.5, .5)), lonv + float(rs.uniform(-.5, .5))
        elev, ele = float(rs.uniform(0, 3000)), float(rs.uniform(0, 4000))
        az0 = float(rs.uniform(-180, 180)); span = float(rs.uniform(20, 359))
        W, H = int(rs.integers(100, 4000)), int(rs.integers(50, 800))
        x, y, r = d(), d(), d()
        args = (latv, np.cos(np.radians(latv)), lonv, elev, lat, lon, ele, np.radians(az0), np.radians(az0 + span), W, H)
        ok = bool(L.horizonator_project(C.byref(x), C.byref(y), C.byref(r), *args))
        out["project"].append(dict(args=list(map(float, args[:-2])) + [W, H], ok=ok,
                                   x=x.value if ok else None, y=y.value if ok else None, range=r.value if ok else None))
    for _ in range(60):
        latv, lonv = float(rs.uniform(-60, 60)), float(rs.uniform(-179, 179))
        az0 = float(rs.uniform(-180, 180)); span = float(rs.uniform(20, 359))
        W, H = int(rs.integers(100, 4000)), int(rs.integers(50, 800))
        px, py = int(rs.integers(0, W)), int(rs.integers(0, H))
        rng = float(rs.uniform(100, 100000))
        which = int(rs.integers(0, 3))
        r_enh, r_en = (rng, -1.) if which == 0 else ((-1., rng) if which == 1 else (rng, rng))
        la, lo = C.c_float(), C.c_float()
        args = (px, py, r_enh, r_en, latv, np.cos(np.radians(latv)), lonv, az0, az0 + span, W, H)
        ok = bool(L.horizonator_unproject(C.byref(la), C.byref(lo), *args))
        out["unproject"].append(dict(args=[px, py] + list(map(float, args[2:-2])) + [W, H], ok=ok,
                                     lat=float(la.value) if ok else None, lon=float(lo.value) if ok else None))
    return out


def renders(tiles):
    out = {}
    for name, W, H, R, az0, az1, zn, zf, znc, zfc, lat, lon in RENDER_SCENES:
        r = binding.Reference(C1_LAT, C1_LON, W, H, dir_dems=tiles, render_radius_cells=R, threads=1)
        kw = {} if lat is None else dict(lat=lat, lon=lon)
        img, rng = r.render(az0, az1, znear=zn, zfar=zf, znear_color=znc, zfar_color=zfc, **kw)
        assert (rng > 0).mean() > 0.01, name
        np.savez_compressed(os.path.join(HERE, "render_%s.npz" % name), image=img, ranges=rng,
                            params=np.array([W, H, R, az0, az1, zn, zf, znc, zfc,
                                             -1000. if lat is None else lat, -1000. if lon is None else lon], np.float64),
                            viewer_z=np.float32(r.viewer_z))
        out[name] = dict(viewer_z_at_init=float(r.viewer_z), hit_fraction=float((rng > 0).mean()))
        binding.Reference.lib().horizonator_deinit(C.byref(r.ctx))
    return out


def move_vectors(tiles):
    """Automatic eye height (lib:775-789) at a few positions inside the C1 square."""
    r = binding.Reference(C1_LAT, C1_LON, 64, 16, dir_dems=tiles, render_radius_cells=600, threads=1)
    rs = np.random.default_rng(3)
    out = [dict(lat=C1_LAT, lon=C1_LON, viewer_z=float(r.viewer_z))]
    for _ in range(25):
        lat = C1_LAT + float(rs.uniform(-0.4, 0.4)); lon = C1_LON + float(rs.uniform(-0.4, 0.4))
        out.append(dict(lat=lat, lon=lon, viewer_z=float(r.move(lat, lon))))
    binding.Reference.lib().horizonator_deinit(C.byref(r.ctx))
    return out


def main():
    if not os.path.isdir("/root/reference"):
        raise SystemExit("needs /root/reference (run in the development container)")
    binding.build(ref=True)
    tmp = tempfile.mkdtemp(prefix="hz_golden_")
    tiles = synth.config1_tiles(os.path.join(tmp, "c1"))
    holes = os.path.join(tmp, "holes")
    synth.write_tiles(holes, (34, 35), (-118, -117), seed=7, skip=((35, -118), (34, -117)))
    open(os.path.join(holes, synth.tile_name(34, -117)), "wb").close()

    def dump(name, obj):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(obj, f, indent=0, separators=(",", ":"))
            f.write("\n")

    dump("dem_geometry.json", dem_geometry(tmp))
    dump("dem_samples.json", dem_samples(tiles, holes))
    dump("geometry.json", geometry_vectors())
    dump("move.json", move_vectors(tiles))
    dump("renders.json", renders(tiles))
    import hashlib
    dump("tiles_sha256.json", {f: hashlib.sha256(open(os.path.join(tiles, f), "rb").read()).hexdigest()
                               for f in sorted(os.listdir(tiles))})
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
