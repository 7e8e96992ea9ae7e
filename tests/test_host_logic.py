"""Host-side logic of libhorizonator that needs no GPU: the DEM layer of include/dem.h (hz_dem.cpp, replaces
/root/reference/dem.c) and the pure projection helpers (replace horizonator-lib.c:1053-1213), checked
bit-for-bit against the golden vectors generated from the reference's own code (tests/golden/make_golden.py)
and, where oracle/_ref is present, against that code directly."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import C1_LAT, C1_LON

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def hz():
    import horizonator_b200
    return horizonator_b200


def test_synthetic_tiles_are_the_golden_ones(tiles_c1):
    """Every golden render/sample vector depends on the generated tiles being the ones the vectors were made on."""
    want = golden("tiles_sha256.json")
    for name, sha in want.items():
        assert hashlib.sha256(open(os.path.join(tiles_c1, name), "rb").read()).hexdigest() == sha, name


def _empty_tiles(tmp_path, lat, lon, tag):
    from tools.synth import tile_name
    d = tmp_path / ("empty_%s" % tag)
    d.mkdir(exist_ok=True)
    for la in range(int(np.floor(lat)) - 3, int(np.floor(lat)) + 4):
        for lo in range(int(np.floor(lon)) - 3, int(np.floor(lon)) + 4):
            (d / tile_name(la, lo)).write_bytes(b"")
    return str(d)


def test_dem_init_geometry_matches_reference(hz, tmp_path):
    """dem.c:78-179: radius in cells, origin tile/cell, tile counts, failure modes."""
    for k, g in enumerate(golden("dem_geometry.json")):
        d = _empty_tiles(tmp_path, g["lat"], g["lon"], str(k))
        ctx = hz.dem_context_t()
        ok = bool(hz.lib.horizonator_dem_init(C.byref(ctx), g["lat"], g["lon"], g["radius_cells"], g["radius_m"],
                                              os.fsencode(d), g["SRTM1"]))
        assert ok == g["ok"], g
        if not ok:
            continue
        assert list(ctx.origin_dem_lon_lat) == g["origin_dem_lon_lat"], g
        assert list(ctx.origin_dem_cellij) == g["origin_dem_cellij"], g
        assert list(ctx.Ndems_ij) == g["Ndems_ij"], g
        assert ctx.radius_cells == g["R"] and ctx.cells_per_deg == g["cells_per_deg"], g
        b = [C.c_float() for _ in range(4)]
        hz.lib.horizonator_dem_bounds_latlon_deg(C.byref(ctx), *[C.byref(x) for x in b])
        assert [float(np.float32(x.value)) for x in b] == g["bounds_lat0_lon0_lat1_lon1"], g
        hz.lib.horizonator_dem_deinit(C.byref(ctx))
        hz.lib.horizonator_dem_deinit(C.byref(ctx))          # idempotent


def test_survey_known_answers(hz, tmp_path):
    """The values SURVEY.md 8d lists for BASELINE configs 1 and 2 (obtained there by running dem.c)."""
    g = {(x["radius_cells"], x["radius_m"], x["SRTM1"]): x for x in golden("dem_geometry.json") if x["ok"]}
    c1 = g[(1200, -1.0, False)]
    assert c1["origin_dem_lon_lat"] == [-118, 34] and c1["origin_dem_cellij"] == [1, 1] and c1["Ndems_ij"] == [2, 2]
    c2 = [x for x in golden("dem_geometry.json") if x["ok"] and x["radius_m"] == 150000.0 and x["lat"] > 30][0]
    assert c2["R"] == 5858 and c2["origin_dem_lon_lat"] == [-119, 32] and c2["origin_dem_cellij"] == [1343, 1343]
    assert c2["Ndems_ij"] == [4, 4]


def _holes_dir(tmp_path, tiles_c1):
    from tools import synth
    d = str(tmp_path / "holes")
    synth.write_tiles(d, (34, 35), (-118, -117), seed=7, skip=((35, -118), (34, -117)))
    open(os.path.join(d, synth.tile_name(34, -117)), "wb").close()
    return d


def test_dem_sample_matches_reference(hz, tiles_c1, tmp_path):
    """dem.c:264-309 incl. the shared tile edge, voids/negatives -> 0, missing and zero-length tiles -> 0."""
    dirs = {"c1": tiles_c1, "c1_small": tiles_c1, "holes": _holes_dir(tmp_path, tiles_c1)}
    for g in golden("dem_samples.json"):
        ctx = hz.dem_context_t()
        assert hz.lib.horizonator_dem_init(C.byref(ctx), C1_LAT, C1_LON, g["R"], -1.0, os.fsencode(dirs[g["name"]]), False)
        got = [int(hz.lib.horizonator_dem_sample(C.byref(ctx), i, j)) for i, j in g["points"]]
        assert got == g["values"], g["name"]
        assert min(got) >= -1 and max(got) > 500
        hz.lib.horizonator_dem_deinit(C.byref(ctx))


def test_dem_init_rejects_wrong_tile_size_and_handles_home(hz, tmp_path, monkeypatch):
    from tools.synth import tile_name
    d = tmp_path / ".horizonator" / "DEMs_SRTM3"
    d.mkdir(parents=True)
    ctx = hz.dem_context_t()
    monkeypatch.setenv("HOME", str(tmp_path))
    # "~/" expansion (dem.c:48-66): all four tiles missing -> warnings, elevation 0, success
    assert hz.lib.horizonator_dem_init(C.byref(ctx), C1_LAT, C1_LON, 8, -1.0, b"~/.horizonator/DEMs_SRTM3", False)
    assert hz.lib.horizonator_dem_sample(C.byref(ctx), 3, 3) == 0
    hz.lib.horizonator_dem_deinit(C.byref(ctx))
    # a file that is neither empty nor (cpd+1)^2*2 bytes (dem.c:234-239)
    (d / tile_name(35, -117)).write_bytes(b"\0" * 1000)
    assert not hz.lib.horizonator_dem_init(C.byref(ctx), C1_LAT + 0.5, C1_LON + 0.5, 8, -1.0, os.fsencode(str(d)), False)
    # an SRTM3-sized file offered as SRTM1
    (d / tile_name(35, -117)).write_bytes(b"\0" * (1201 * 1201 * 2))
    assert hz.lib.horizonator_dem_init(C.byref(ctx), C1_LAT + 0.5, C1_LON + 0.5, 8, -1.0, os.fsencode(str(d)), False)
    hz.lib.horizonator_dem_deinit(C.byref(ctx))
    assert not hz.lib.horizonator_dem_init(C.byref(ctx), C1_LAT + 0.5, C1_LON + 0.5, 8, -1.0, os.fsencode(str(d)), True)
    monkeypatch.delenv("HOME")
    assert not hz.lib.horizonator_dem_init(C.byref(ctx), C1_LAT, C1_LON, 8, -1.0, b"~/.horizonator/DEMs_SRTM3", False)


def test_projection_helpers_match_reference(hz):
    """horizonator_x_from_az / _project / _unproject vs the reference's own (golden, bit-exact: same double/float
    expressions, horizonator-lib.c:1062-1213)."""
    g = golden("geometry.json")
    d = C.c_double
    for v in g["x_from_az"]:
        x, per = d(), d()
        ok = bool(hz.lib.horizonator_x_from_az(C.byref(x), C.byref(per), v["az_rad"], v["az_rad0"], v["az_rad1"], v["width"]))
        assert ok == v["ok"], v
        if ok:
            assert x.value == v["x"] and per.value == v["per"], v
    for v in g["project"]:
        x, y, r = d(), d(), d()
        ok = bool(hz.lib.horizonator_project(C.byref(x), C.byref(y), C.byref(r), *v["args"]))
        assert ok == v["ok"], v
        if ok:
            assert (x.value, y.value, r.value) == (v["x"], v["y"], v["range"]), v
    for v in g["unproject"]:
        la, lo = C.c_float(), C.c_float()
        ok = bool(hz.lib.horizonator_unproject(C.byref(la), C.byref(lo), *v["args"]))
        assert ok == v["ok"], v
        if ok:
            assert (float(la.value), float(lo.value)) == (v["lat"], v["lon"]), v
    assert any(v["ok"] for v in g["project"]) and any(not v["ok"] for v in g["project"])


def test_project_unproject_round_trip(hz):
    """project -> unproject returns to the start (float32 lat/lon).  The window is
    359.9 degrees wide: an exactly 360-degree one makes the reference's double-precision helper return false
    (SURVEY appendix B, Q5), and so does this one (checked at the end)."""
    d = C.c_double
    rs = np.random.default_rng(1)
    n_ok = 0
    for _ in range(200):
        latv, lonv = 34.0, -117.0
        lat, lon = latv + rs.uniform(-.3, .3), lonv + rs.uniform(-.3, .3)
        W, H = 3600, 600
        x, y, r = d(), d(), d()
        if not hz.lib.horizonator_project(C.byref(x), C.byref(y), C.byref(r), latv, np.cos(np.radians(latv)), lonv, 1000.,
                                          lat, lon, 1200., np.radians(-180.0), np.radians(179.9), W, H):
            continue
        px, py = int(round(x.value)), int(round(y.value))
        la, lo = C.c_float(), C.c_float()
        assert hz.lib.horizonator_unproject(C.byref(la), C.byref(lo), px, py, r.value, -1., latv,
                                            np.cos(np.radians(latv)), lonv, -180.0, 179.9, W, H)
        # half a pixel of 0.1 degrees at up to ~40 km, plus float32 lat/lon
        tol = np.degrees(np.radians(0.12) * r.value / 6371000.0) + 2e-5
        assert abs(la.value - lat) < tol and abs(lo.value - lon) < tol / np.cos(np.radians(latv)) + 2e-5
        n_ok += 1
    assert n_ok > 100
    x, per = d(), d()
    assert not hz.lib.horizonator_x_from_az(C.byref(x), C.byref(per), 0.3, -np.pi, np.pi, 3600)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhorizonator_ref.so")),
                    reason="oracle/_ref not built on this machine")
def test_dem_layer_against_reference_build_directly(hz, tiles_c1):
    """Scattered samples of the whole C1 mosaic: product's hz_dem.cpp vs the reference's dem.c (oracle/_ref)."""
    from oracle import binding
    L = binding.Reference.lib()
    a, b = hz.dem_context_t(), hz.dem_context_t()
    for R in (7, 600, 1200):
        assert hz.lib.horizonator_dem_init(C.byref(a), C1_LAT, C1_LON, R, -1.0, os.fsencode(tiles_c1), False)
        assert L.horizonator_dem_init(C.byref(b), C1_LAT, C1_LON, R, -1.0, os.fsencode(tiles_c1), False)
        assert bytes(a)[320:] == bytes(b)[320:]              # every int field of the struct
        rs = np.random.default_rng(R)
        for _ in range(3000):
            i, j = int(rs.integers(0, 2 * R)), int(rs.integers(0, 2 * R))
            assert hz.lib.horizonator_dem_sample(C.byref(a), i, j) == L.horizonator_dem_sample(C.byref(b), i, j)
        hz.lib.horizonator_dem_deinit(C.byref(a))
        L.horizonator_dem_deinit(C.byref(b))


def test_pinned_pool_recycles_blocks_with_array_lifetime(hz, monkeypatch):
    """render() hands out arrays built on pooled page-locked blocks; a block returns to the pool when the last view
    of the array is gone.  The allocator is faked here (no CUDA on this machine)."""
    import gc

    class FakeBlock:
        made = 0

        def __init__(self, nbytes):
            self.mem = (C.c_uint8 * nbytes)()
            self.ptr = C.addressof(self.mem)
            FakeBlock.made += 1

    monkeypatch.setattr(hz, "_PinnedBlock", FakeBlock)
    pool = hz._PinnedPool(keep=100)
    a = pool.array((4, 5, 3), np.uint8)
    a[:] = 7
    assert a.flags.writeable and a.shape == (4, 5, 3) and int(a.sum()) == 7 * 60
    view = a[1:3]
    del a
    gc.collect()
    assert pool.kept == 0                       # still leased through the view
    del view
    gc.collect()
    assert pool.kept == 60
    b = pool.array((4, 5, 3), np.uint8)         # the same block again
    assert FakeBlock.made == 1 and pool.kept == 0
    c = pool.array((10, 10), np.float32)        # 400 bytes: more than the pool keeps
    del b, c
    gc.collect()
    assert pool.kept == 60 and FakeBlock.made == 2


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhorizonator_ref.so")),
                    reason="oracle/_ref not built on this machine")
def test_projection_helpers_against_reference_build_directly(hz):
    """Differential run of horizonator_x_from_az / _project / _unproject against the reference's own compiled code
    (oracle/_ref) on a few thousand random inputs, windows across the +-180 seam and wider than a circle included:
    same return flag and bit-identical outputs."""
    from oracle import binding
    L = binding.Reference.lib()
    d = C.c_double
    rs = np.random.default_rng(42)
    n_true = n_false = 0
    for _ in range(3000):
        az0 = float(rs.uniform(-4 * np.pi, 4 * np.pi))
        az1 = az0 + float(rs.choice([rs.uniform(0.01, 2 * np.pi), 2 * np.pi, rs.uniform(2 * np.pi, 4 * np.pi), -rs.uniform(0.01, 3.)]))
        az = float(rs.uniform(-4 * np.pi, 4 * np.pi))
        W = int(rs.integers(1, 40000))
        xa, pa, xb, pb = d(), d(), d(), d()
        oka = bool(hz.lib.horizonator_x_from_az(C.byref(xa), C.byref(pa), az, az0, az1, W))
        okb = bool(L.horizonator_x_from_az(C.byref(xb), C.byref(pb), az, az0, az1, W))
        assert oka == okb, (az, az0, az1, W)
        if oka:
            assert xa.value == xb.value and pa.value == pb.value, (az, az0, az1, W)
        n_true += oka; n_false += not oka
    assert n_true > 500 and n_false > 50

    n_true = n_false = 0
    for _ in range(3000):
        latv, lonv = float(rs.uniform(-60., 60.)), float(rs.uniform(-179., 179.))
        coslat = float(np.cos(np.radians(latv)))
        lat, lon = latv + float(rs.uniform(-.8, .8)), lonv + float(rs.uniform(-.8, .8))
        az0 = float(rs.uniform(-2 * np.pi, 2 * np.pi)); az1 = az0 + float(rs.uniform(0.02, 2 * np.pi - 1e-3))
        W, H = int(rs.integers(16, 40000)), int(rs.integers(8, 5000))
        args = (latv, coslat, lonv, float(rs.uniform(0., 4000.)), lat, lon, float(rs.uniform(-100., 5000.)), az0, az1, W, H)
        a, b = [d(), d(), d()], [d(), d(), d()]
        oka = bool(hz.lib.horizonator_project(*[C.byref(v) for v in a], *args))
        okb = bool(L.horizonator_project(*[C.byref(v) for v in b], *args))
        assert oka == okb, args
        if oka:
            assert [v.value for v in a] == [v.value for v in b], args
        n_true += oka; n_false += not oka

        px, py = int(rs.integers(0, W)), int(rs.integers(0, H))
        which = int(rs.integers(0, 3)); r = float(rs.uniform(10., 150000.))
        r_enh, r_en = (r, -1.) if which == 0 else ((-1., r) if which == 1 else (r, r))
        uargs = (px, py, r_enh, r_en, latv, coslat, lonv, float(np.degrees(az0)), float(np.degrees(az1)), W, H)
        la, lo, lb, lob = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        oka = bool(hz.lib.horizonator_unproject(C.byref(la), C.byref(lo), *uargs))
        okb = bool(L.horizonator_unproject(C.byref(lb), C.byref(lob), *uargs))
        assert oka == okb, uargs
        if oka:
            assert la.value == lb.value and lo.value == lob.value, uargs
    assert n_true > 300 and n_false > 300


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhorizonator_ref.so")),
                    reason="oracle/_ref not built on this machine")
def test_dem_init_geometry_against_reference_build_on_random_viewpoints(hz, tmp_path):
    """Tile selection is float/int arithmetic that must come out the same as dem.c:78-243 everywhere on the globe:
    random viewpoints (both hemispheres, near the date line, near tile edges), radii in cells or metres, SRTM3 and
    SRTM1, against the reference's compiled dem.c.  The tile directory is empty (all tiles read as sea level), so
    only the geometry is compared: success flag and every integer field of the context."""
    from oracle import binding
    L = binding.Reference.lib()
    empty = os.fsencode(str(tmp_path))
    rs = np.random.default_rng(9)
    n_ok = n_fail = 0
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)                      # both libraries warn about every missing tile
    try:
        for k in range(400):
            lat = float(rs.uniform(-59., 59.)); lon = float(rs.uniform(-179.5, 179.5))
            if k % 5 == 0:                   # on or next to a tile corner
                lat = float(np.round(lat)) + float(rs.choice([0., 1e-4, -1e-4, 1. / 2400.]))
                lon = float(np.round(lon)) + float(rs.choice([0., 1e-4, -1e-4, 1. / 2400.]))
            srtm1 = bool(rs.random() < 0.4)
            if rs.random() < 0.5:
                rc, rm = int(rs.integers(1, 7300 if srtm1 else 2500)), -1.0
            else:
                rc, rm = -1, float(rs.uniform(500., 260000.))
            a, b = hz.dem_context_t(), hz.dem_context_t()
            oka = bool(hz.lib.horizonator_dem_init(C.byref(a), lat, lon, rc, rm, empty, srtm1))
            okb = bool(L.horizonator_dem_init(C.byref(b), lat, lon, rc, rm, empty, srtm1))
            assert oka == okb, (lat, lon, rc, rm, srtm1)
            if oka:
                assert bytes(a)[320:] == bytes(b)[320:], (lat, lon, rc, rm, srtm1)
                hz.lib.horizonator_dem_deinit(C.byref(a))
                L.horizonator_dem_deinit(C.byref(b))
            n_ok += oka; n_fail += not oka
    finally:
        os.dup2(saved, 2); os.close(saved); os.close(devnull)
    assert n_ok > 100 and n_fail > 5, (n_ok, n_fail)
