"""The oracle's GL rules pinned against a REAL OpenGL implementation.

tests/golden/llvmpipe_*.npz are renders of the reference's own horizonator-lib.c + dem.c, compiled unmodified, on
Mesa's llvmpipe (tests/golden/make_golden_llvmpipe.py; oracle/mesa/ explains how a GL context exists in an image
without X).  Here the CPU oracle -- whose rasterisation rules F1-F9 restate the OpenGL specification -- is held
against them at the north_star tolerances, and so are the older fixtures that came from the fake-GL build of the
reference.  Where the Mesa build is present (this container, and any box with the same image) the reference is also
run live on llvmpipe, in a child process.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from compare import compare_renders

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
C1_LAT, C1_LON = 35.0 + 1.0 / 2400.0, -117.0 + 1.0 / 2400.0

SCENES = sorted(f[len("llvmpipe_"):-len(".npz")] for f in os.listdir(GOLDEN) if f.startswith("llvmpipe_") and f.endswith(".npz"))


def _load(name):
    g = np.load(os.path.join(GOLDEN, "llvmpipe_%s.npz" % name))
    return g, [float(x) for x in g["params"]]


def test_fixture_set():
    meta = json.load(open(os.path.join(GOLDEN, "llvmpipe.json")))
    assert "llvmpipe" in meta["gl_renderer"] and "Mesa" in meta["gl_version"]
    assert sorted(k for k in meta["scenes"] if not k.startswith("fullsize_")) == SCENES and len(SCENES) >= 7
    assert {"fullsize_c1", "fullsize_c2"} <= set(meta["scenes"])
    # horizonator_move()'s automatic eye heights (a render + read-back each) came out the same on llvmpipe as on the
    # fake GL: the generator asserts it vector by vector and records how many it checked
    assert meta["move_json_reproduced"] == len(json.load(open(os.path.join(GOLDEN, "move.json"))))


@pytest.mark.parametrize("name", SCENES)
def test_oracle_agrees_with_llvmpipe(tiles_c1, name):
    from oracle.binding import Oracle
    g, (W, H, R, az0, az1, zn, zf, znc, zfc, lat, lon, vz) = _load(name)
    o = Oracle(C1_LAT, C1_LON, int(W), int(H), dir_dems=tiles_c1, render_radius_cells=int(R),
               viewer_z=None if vz < 0 else vz, threads=min(8, os.cpu_count() or 1))
    assert np.float32(o.viewer_z) == g["viewer_z"]
    kw = {} if lat <= -1000. else dict(lat=lat, lon=lon)
    img, rng = o.render(az0, az1, znear=zn, zfar=zf, znear_color=znc, zfar_color=zfc, **kw)
    s = compare_renders(img, rng, g["image"], g["ranges"])
    print("oracle vs llvmpipe", name, s)
    assert s["ok"], s
    # far tighter than the north_star bar: at most a handful of pixels see terrain in one and sky in the other, and
    # every range disagreement sits on a silhouette
    assert ((rng > 0) != (g["ranges"] > 0)).sum() <= max(2, rng.size // 20000), s
    assert s["off_silhouette"] == 0, s
    assert s["agreement"] >= 0.998, s


@pytest.mark.parametrize("name", ["circle_small", "quarter", "seam_odd_h", "moved"])
def test_fake_gl_fixtures_agree_with_llvmpipe(name):
    """render_*.npz (reference on the fake GL) against llvmpipe_*.npz (reference on Mesa): same scenes."""
    f = np.load(os.path.join(GOLDEN, "render_%s.npz" % name))
    g, p = _load(name)
    assert list(f["params"]) == p[:11] and f["viewer_z"] == g["viewer_z"]
    s = compare_renders(f["image"], f["ranges"], g["image"], g["ranges"])
    assert s["ok"] and s["coverage_agreement"] == 1.0 and s["off_silhouette"] == 0, s


C2_LAT, C2_LON = 34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0


@pytest.mark.parametrize("which", ["c1", "c2"])
def test_oracle_agrees_with_llvmpipe_at_full_size(tiles_c1, which):
    """BASELINE configs[0] (3600x300, 2x2 SRTM3 tiles, 11.5 M triangles) and configs[1] (the benchmark panorama:
    3600x600, 150 km of SRTM1, 274 M triangles): the oracle against the reference's render on llvmpipe."""
    from oracle.binding import Oracle
    from tools import synth
    g = np.load(os.path.join(GOLDEN, "fullsize_%s_llvmpipe.npz" % which))
    threads = os.cpu_count() or 1
    if which == "c1":
        o = Oracle(C1_LAT, C1_LON, 3600, 300, dir_dems=tiles_c1, render_radius_cells=1200, threads=threads)
        zfar = 100000.
    else:
        tiles = synth.config2_tiles(os.environ.get("HZ_BENCH_TILES", "/tmp/hz_tiles_c2"))
        o = Oracle(C2_LAT, C2_LON, 3600, 600, SRTM1=True, dir_dems=tiles, render_radius_m=150000., threads=threads)
        zfar = 150000.
    assert np.float32(o.viewer_z) == g["viewer_z"]
    img, rng = o.render(-180.05, 179.95, znear=100., zfar=zfar)
    s = compare_renders(img, rng, g["image"], g["ranges"])
    print("oracle vs llvmpipe, full size", which, s)
    assert s["ok"], s
    assert s["coverage_agreement"] >= 0.99999 and s["agreement"] >= 0.9995 and s["off_silhouette"] == 0, s


WORKER = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
from oracle import binding
from compare import compare_renders
out = {}
for name in %(scenes)r:
    g = np.load(os.path.join(%(golden)r, "llvmpipe_%%s.npz" %% name))
    W, H, R, az0, az1, zn, zf, znc, zfc, lat, lon, vz = (float(x) for x in g["params"])
    r = binding.MesaReference(%(lat)r, %(lon)r, int(W), int(H), dir_dems=%(tiles)r, render_radius_cells=int(R),
                              viewer_z=None if vz < 0 else vz, threads=2)
    kw = {} if lat <= -1000. else dict(lat=lat, lon=lon)
    img, rng = r.render(az0, az1, znear=zn, zfar=zf, znear_color=znc, zfar_color=zfc, **kw)
    version, renderer = r.gl_strings()
    r.close()
    s = compare_renders(img, rng, g["image"], g["ranges"])
    s["identical"] = bool(np.array_equal(img, g["image"]) and np.array_equal(rng, g["ranges"]))
    s["renderer"] = renderer
    out[name] = s
print("RESULT " + json.dumps(out))
"""


def test_reference_runs_live_on_llvmpipe(tiles_c1):
    """The reference, unmodified, on Mesa llvmpipe in a child process (a driver crash must not take pytest along);
    what it renders now equals what was recorded (bit for bit on the same CPU family; within the parity bar anywhere,
    since llvmpipe compiles its shaders for the host CPU)."""
    from oracle import binding
    if not binding.have_mesa():
        pytest.skip("oracle/_ref/libhorizonator_mesa.so or the image's Mesa libGL is absent")
    code = WORKER % dict(root=ROOT, tests=HERE, golden=GOLDEN, scenes=["circle_small", "quarter", "high_eye"],
                         lat=C1_LAT, lon=C1_LON, tiles=tiles_c1)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    print(res)
    for name, s in res.items():
        assert "llvmpipe" in s["renderer"]
        # not s["ok"]: that also demands the reference's sky colour of the image under test, and llvmpipe itself
        # leaves the odd far-plane pixel black with depth 1.0 ("what llvmpipe does differently")
        assert s["coverage_agreement"] >= 0.9999 and s["agreement"] >= 0.999 and s["off_silhouette"] == 0, (name, s)
        assert s["red_max_diff_where_agree"] <= 1, (name, s)


def test_random_scenes_oracle_vs_live_llvmpipe():
    """tools/llvmpipe_sweep.py: its 15 fixed edge cases, then a fresh seed of random windows (full circles, zooms, windows across the +-180 seam),
    sizes, radii, eye positions and heights, depth ranges, tiles with holes -- the oracle against the reference running
    live on llvmpipe.  (The edge cases + 60 scenes of seed 1 are recorded in profiles/r01C_oracle_vs_llvmpipe_sweep.json.)"""
    from oracle import binding
    if not binding.have_mesa():
        pytest.skip("oracle/_ref/libhorizonator_mesa.so or the image's Mesa libGL is absent")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "llvmpipe_sweep.py"), "--scenes", "10", "--seed", "5"],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    tot = json.loads(p.stdout.strip().splitlines()[-1])
    print(tot)
    assert tot["not_ok"] == 0 and tot["off_silhouette"] == 0 and tot["eye_height_differs"] == 0, tot
    assert tot["coverage_agreement_overall"] >= 0.9999 and tot["worst_agreement"] >= 0.995, tot
    assert tot["terrain_pixels"] > 10000


def test_pick_on_fake_gl_equals_pick_on_llvmpipe(tiles_c1):
    """horizonator_pick() (a one-pixel depth read-back + unproject, horizonator-lib.c:1216-1296) of the reference on
    the oracle's GL restatement against the same calls recorded from the reference on llvmpipe."""
    import ctypes as C
    from oracle import binding
    if not binding.have_ref():
        pytest.skip("oracle/_ref not built")
    pick = json.load(open(os.path.join(GOLDEN, "llvmpipe.json")))["pick"]
    W, H, R, az0, az1, zn, zf = pick["scene"]
    r = binding.Reference(C1_LAT, C1_LON, int(W), int(H), dir_dems=tiles_c1, render_radius_cells=int(R), threads=2)
    try:
        r.render(az0, az1, znear=zn, zfar=zf)
        for p in pick["points"]:
            la, lo = C.c_float(), C.c_float()
            ok = bool(r.lib().horizonator_pick(C.byref(r.ctx), C.byref(la), C.byref(lo), p["x"], p["y"]))
            assert ok == p["ok"], p
            if ok:
                assert la.value == np.float32(p["lat"]) and lo.value == np.float32(p["lon"]), (p, la.value, lo.value)
    finally:
        r.close()
