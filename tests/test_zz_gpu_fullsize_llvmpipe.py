"""The CUDA path against the reference's own renders on Mesa llvmpipe at the two BASELINE workloads, full size
(tests/golden/fullsize_c{1,2}_llvmpipe.npz, made by tests/golden/make_golden_llvmpipe.py): configs[0] -- 3600x300 over
2x2 SRTM3 tiles -- and configs[1], the benchmark panorama -- 3600x600 over 150 km of SRTM1, 274 M triangles.  No oracle
code is involved: this is the north_star parity statement itself ("checked against the reference's own GL path
(Mesa llvmpipe offscreen) on the same synthetic .hgt inputs")."""
import os

import numpy as np
import pytest

from compare import compare_renders

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
C1_LAT, C1_LON = 35.0 + 1.0 / 2400.0, -117.0 + 1.0 / 2400.0
C2_LAT, C2_LON = 34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0


def test_config1_full_size_matches_llvmpipe(tiles_c1):
    import horizonator_b200 as hz
    g = np.load(os.path.join(GOLDEN, "fullsize_c1_llvmpipe.npz"))
    h = hz.horizonator(C1_LAT, C1_LON, 3600, 300, dir_dems=tiles_c1, render_radius_cells=1200)
    img, rng = h.render(-180.05, 179.95, znear=100., zfar=100000.)
    s = compare_renders(img, rng, g["image"], g["ranges"])
    print("CUDA vs llvmpipe, config 1 full size", s)
    assert s["ok"], s
    assert s["coverage_agreement"] >= 0.9995, s


def test_config2_full_size_matches_llvmpipe():
    import horizonator_b200 as hz
    from tools import synth
    tiles = synth.config2_tiles(os.environ.get("HZ_BENCH_TILES", "/tmp/hz_tiles_c2"))
    g = np.load(os.path.join(GOLDEN, "fullsize_c2_llvmpipe.npz"))
    h = hz.horizonator(C2_LAT, C2_LON, 3600, 600, SRTM1=True, dir_dems=tiles, render_radius_m=150000.)
    img, rng = h.render(-180.05, 179.95, znear=100., zfar=150000.)
    s = compare_renders(img, rng, g["image"], g["ranges"])
    print("CUDA vs llvmpipe, config 2 full size", s)
    assert s["ok"], s
    assert s["coverage_agreement"] >= 0.9995, s


def test_pick_matches_llvmpipe(tiles_c1):
    """horizonator_pick() of the product against the picks recorded from the reference on llvmpipe (the reference on
    the oracle's GL restatement reproduces them bit for bit, tests/test_llvmpipe.py)."""
    import ctypes as C
    import json
    import horizonator_b200 as hz
    pick = json.load(open(os.path.join(GOLDEN, "llvmpipe.json")))["pick"]
    W, H, R, az0, az1, zn, zf = pick["scene"]
    h = hz.horizonator(C1_LAT, C1_LON, int(W), int(H), dir_dems=tiles_c1, render_radius_cells=int(R))
    h.render(az0, az1, znear=zn, zfar=zf)
    close = flags_differ = 0
    for p in pick["points"]:
        la, lo = C.c_float(), C.c_float()
        ok = bool(hz.lib.horizonator_pick(C.byref(h.context), C.byref(la), C.byref(lo), p["x"], p["y"]))
        flags_differ += ok != p["ok"]
        if ok and p["ok"] and abs(la.value - p["lat"]) <= 2e-5 and abs(lo.value - p["lon"]) <= 2e-5:
            close += 1
    n_ok = sum(1 for p in pick["points"] if p["ok"])
    print("picks within 2e-5 degrees of llvmpipe's:", close, "of", n_ok, "; hit/miss differs on", flags_differ)
    assert flags_differ <= 1 and close >= n_ok - 3      # a silhouette pixel may see the neighbouring surface


def test_random_scenes_match_live_llvmpipe(tiles_c1, tiles_holes, tmp_path):
    """Random views (tools/llvmpipe_sweep.py, seed 3: full circles, zooms, windows across the seam, eye heights, depth
    ranges, tiles with holes) rendered by the reference on llvmpipe in a child process, right here on the box's host
    cores, and by the CUDA path; compared at the north_star tolerances.  No oracle result is used."""
    import ctypes as C
    import json
    import subprocess
    import sys
    import horizonator_b200 as hz
    from oracle import binding                      # only to ask whether the llvmpipe build exists
    if not binding.have_mesa():
        pytest.skip("oracle/_ref/libhorizonator_mesa.so or the image's Mesa libGL is absent")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dump = str(tmp_path / "llvmpipe")
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "llvmpipe_sweep.py"), "--scenes", "12", "--seed", "3",
                        "--no-edge-cases", "--dump", dump], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    scenes = json.load(open(os.path.join(dump, "scenes.json")))["scenes"]
    assert len(scenes) == 12
    worst, terrain = 1.0, 0
    for k, sc in enumerate(scenes):
        g = np.load(os.path.join(dump, "scene_%03d.npz" % k))
        ctx = hz.context_t()
        z = C.c_float(-1. if sc["viewer_z"] is None else sc["viewer_z"])
        tiles = tiles_holes if sc["holes"] else tiles_c1
        assert hz.lib.horizonator_init(C.byref(ctx), C1_LAT, C1_LON, C.byref(z), sc["W"], sc["H"], sc["R"], -1.0, True, False,
                                       False, os.fsencode(tiles), None, None, None, False)
        try:
            assert np.float32(z.value) == g["viewer_z"], sc
            assert hz.lib.horizonator_pan_zoom(C.byref(ctx), sc["az0"], sc["az1"])
            if sc["lat"] is not None:
                assert hz.lib.horizonator_move(C.byref(ctx), None, sc["lat"], sc["lon"])
            assert hz.lib.horizonator_set_zextents(C.byref(ctx), sc["znear"], sc["zfar"], sc["znear_color"], sc["zfar_color"])
            img = np.empty((sc["H"], sc["W"], 3), np.uint8)
            rng = np.empty((sc["H"], sc["W"]), np.float32)
            assert hz.lib.horizonator_render_offscreen(C.byref(ctx), img.ctypes.data, rng.ctypes.data)
        finally:
            hz.lib.horizonator_deinit(C.byref(ctx))
        s = compare_renders(img, rng, g["image"], g["ranges"])
        print("CUDA vs live llvmpipe", k, sc["kind"], "%dx%d" % (sc["W"], sc["H"]), s)
        assert s["ok"], (sc, s)
        worst = min(worst, s["agreement"]); terrain += int((g["ranges"] > 0).sum())
    assert terrain > 10000
    print("worst agreement over the random scenes:", worst)


def test_config2_batch_of_16_equals_single_renders():
    """The configuration bench.py times -- 16 benchmark-size panoramas per call on the render lanes, with the smaller
    grids views of a batch get -- gives, view by view, exactly what single renders give."""
    import horizonator_b200 as hz
    from tools import synth
    tiles = synth.config2_tiles(os.environ.get("HZ_BENCH_TILES", "/tmp/hz_tiles_c2"))
    W, H = 3600, 600
    h = hz.horizonator(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles, render_radius_m=150000.)
    h.set_zextents(100., 150000.)
    views = [(C2_LAT, C2_LON, -180.05, 179.95)] * 13 + [(C2_LAT + 0.05 * k, C2_LON - 0.04 * k, -180.05, 179.95) for k in (1, 2, 3)]
    bi, br = h.render_batch(views)
    singles = {}
    for k, v in enumerate(views):
        if v not in singles:
            singles[v] = h.render(v[2], v[3], lat=v[0], lon=v[1], znear=100., zfar=150000.)
        i1, r1 = singles[v]
        assert np.array_equal(i1, bi[k]) and np.array_equal(r1, br[k]), k
    g = np.load(os.path.join(GOLDEN, "fullsize_c2_llvmpipe.npz"))
    s = compare_renders(bi[0], br[0], g["image"], g["ranges"])
    print("batch view 0 vs llvmpipe, config 2 full size", s)
    assert s["ok"], s
