"""GPU parity at BASELINE sizes: config 2 (SRTM1, R=5858, 3600x600) against the oracle at the benchmark viewpoint
and at a moved one, a high-resolution window (config 4 flavour) and a narrow zoom window (config 3 flavour).
The oracle needs ~1 s per full-size render on 16 cores; the SRTM1 tiles are generated once per session."""
import os

import numpy as np
import pytest

from compare import compare_renders

pytestmark = pytest.mark.gpu

C2_LAT, C2_LON = 34.0 + 1.0 / 7200.0, -117.0 + 1.0 / 7200.0


@pytest.fixture(scope="module")
def tiles_c2():
    from tools import synth
    return synth.config2_tiles(os.environ.get("HZ_BENCH_TILES", "/tmp/hz_tiles_c2"))


@pytest.fixture(scope="module")
def pair_c2(tiles_c2):
    import horizonator_b200 as hz
    from oracle.binding import Oracle
    W, H = 3600, 600
    h = hz.horizonator(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles_c2, render_radius_m=150000.)
    o = Oracle(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles_c2, render_radius_m=150000., threads=os.cpu_count() or 1)
    return h, o


def test_config2_mosaic_bit_exact_against_the_reference_sampler(pair_c2, tiles_c2):
    """K1 at BASELINE configs[1] size, all 137 M cells, against dem.c itself (compiled unmodified, oracle/_ref):
    horizonator_dem_sample() once per cell, as horizonator-lib.c:435-439 calls it."""
    from oracle import binding
    if not binding.have_ref():
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    h, _ = pair_c2
    rd = binding.ReferenceDem(C2_LAT, C2_LON, SRTM1=True, dir_dems=tiles_c2, render_radius_m=150000., threads=os.cpu_count() or 1)
    want = rd.mosaic()
    rd.close()
    got = h.mosaic()
    assert got.shape == want.shape == (11716, 11716)
    assert np.array_equal(got, want)
    assert int(want.max()) > 1000 and int((want == 0).sum()) > 0        # real terrain, and voids clamped to 0


def test_config2_mosaic_bit_exact_on_a_sample(pair_c2):
    h, o = pair_c2
    m = h.mosaic()
    assert m.shape == (11716, 11716)
    rs = np.random.default_rng(2)
    for _ in range(4000):
        i, j = int(rs.integers(0, 11716)), int(rs.integers(0, 11716))
        assert m[j, i] == o.dem_sample(i, j)
    # every tile boundary row/column of the 4x4 block (dem.c:287-291)
    for g in (3600, 7200, 10800):
        k = g - 1343
        for d in (-1, 0, 1):
            for t in range(0, 11716, 487):
                assert m[k + d, t] == o.dem_sample(t, k + d) and m[t, k + d] == o.dem_sample(k + d, t)


@pytest.mark.parametrize("view", [
    ("benchmark", None, None, -180.05, 179.95),
    ("moved_ne", C2_LAT + 0.31, C2_LON + 0.22, -180.05, 179.95),
    ("moved_sw_quarter", C2_LAT - 0.4, C2_LON - 0.35, 10.0, 100.0),
    ("zoom_10deg", None, None, 40.0, 50.0),
], ids=lambda v: v[0])
def test_config2_full_size_matches_oracle(pair_c2, view):
    h, o = pair_c2
    name, lat, lon, az0, az1 = view
    kw = {} if lat is None else dict(lat=lat, lon=lon)
    if lat is None:
        kw = dict(lat=C2_LAT, lon=C2_LON)          # the pair is shared: always say where the eye is
    img, rng = h.render(az0, az1, znear=100., zfar=150000., **kw)
    img_o, rng_o = o.render(az0, az1, znear=100., zfar=150000., **kw)
    s = compare_renders(img, rng, img_o, rng_o)
    print(name, s)
    assert s["hit_fraction_ref"] > 0.005
    assert s["ok"], s
    # a second render of the same view is bit-identical (the culling depends on timing, the image must not)
    img2, rng2 = h.render(az0, az1, znear=100., zfar=150000., **kw)
    assert np.array_equal(img, img2) and np.array_equal(rng, rng2)


def test_high_resolution_window_and_wedges(tiles_c2):
    """9000x1000 (0.04 degrees per pixel, config 4 flavour) over a 600-cell radius: full render vs the oracle, and
    the same image assembled from 3 azimuth wedges, bit for bit."""
    import torch
    import horizonator_b200 as hz
    from oracle.binding import Oracle
    W, H, R = 9000, 1000, 600
    h = hz.horizonator(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles_c2, render_radius_cells=R)
    img, rng = h.render(-180.02, 179.98, znear=50., zfar=30000.)
    o = Oracle(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles_c2, render_radius_cells=R, threads=4)
    img_o, rng_o = o.render(-180.02, 179.98, znear=50., zfar=30000.)
    s = compare_renders(img, rng, img_o, rng_o)
    print("hires", s, h.last_render_stats())
    assert s["ok"], s
    edges = [0, 3000, 6001, 9000]
    out_i, out_r = np.empty_like(img), np.empty_like(rng)
    for g in range(3):
        x0, x1 = edges[g], edges[g + 1]
        di = torch.empty((H, x1 - x0, 3), dtype=torch.uint8, device="cuda")
        dr = torch.empty((H, x1 - x0), dtype=torch.float32, device="cuda")
        h.render_wedge_device(x0, x1, di.data_ptr(), dr.data_ptr())
        torch.cuda.synchronize()
        out_i[:, x0:x1] = di.cpu().numpy()
        out_r[:, x0:x1] = dr.cpu().numpy()
    assert np.array_equal(out_i, img) and np.array_equal(out_r, rng)


def test_horizon_profile_kernel_matches_reference_implementation(pair_c2):
    import torch
    from horizonator_b200 import sharding
    h, _ = pair_c2
    views = [(C2_LAT + 0.05 * k, C2_LON - 0.04 * k, -180.05, 179.95) for k in range(3)]
    W, H = h.width, h.height
    h.set_zextents(100., 150000.)
    d_rng = torch.empty((3, H, W), dtype=torch.float32, device="cuda")
    h.render_batch_device(views, 0, d_rng.data_ptr(), torch.cuda.current_stream().cuda_stream)
    rows = torch.empty((3, W), dtype=torch.int32, device="cuda")
    top = torch.empty((3, W), dtype=torch.float32, device="cuda")
    h.horizon_profile_device(d_rng.data_ptr(), 3, rows.data_ptr(), top.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want_rows, want_top = sharding.horizon_profile(d_rng)
    assert torch.equal(rows, want_rows) and torch.equal(top, want_top)
    assert (rows >= 0).float().mean() > 0.9        # nearly every column sees terrain somewhere


def test_very_wide_panorama_matches_oracle(tiles_c2):
    """36000 columns (0.01 degrees per pixel, BASELINE config 4's width): window coordinates reach 36000, where a float
    ulp is 1/256 pixel -- the margins of the conservative culling boxes have to allow for that."""
    import horizonator_b200 as hz
    from oracle.binding import Oracle
    W, H, R = 36000, 400, 500
    h = hz.horizonator(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles_c2, render_radius_cells=R)
    img, rng = h.render(-180.005, 179.995, znear=50., zfar=30000.)
    o = Oracle(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles_c2, render_radius_cells=R, threads=8)
    img_o, rng_o = o.render(-180.005, 179.995, znear=50., zfar=30000.)
    s = compare_renders(img, rng, img_o, rng_o)
    print("very wide", s)
    assert s["hit_fraction_ref"] > 0.01
    assert s["ok"], s


def _compare_in_column_chunks(img, rng, img_o, rng_o, chunk=3600):
    """compare_renders() over column chunks (its float64 temporaries for 144 M pixels at once would need ~10 GB)."""
    worst, total_bad, hit = 1.0, 0, 0.0
    W = rng.shape[1]
    for x0 in range(0, W, chunk):
        sl = slice(x0, min(x0 + chunk, W))
        s = compare_renders(img[:, sl], rng[:, sl], img_o[:, sl], rng_o[:, sl])
        assert s["ok"], (x0, s)
        worst = min(worst, s["agreement"]); total_bad += s["off_silhouette"]; hit += s["hit_fraction_ref"] * (sl.stop - sl.start) / W
    return dict(worst_chunk_agreement=worst, off_silhouette=total_bad, hit_fraction_ref=hit)


def test_config4_at_spec_matches_oracle_and_wedges_assemble(tiles_c2):
    """BASELINE configs[3] at its full size: 36000 x 4000 over the 150 km DEM (R = 5858).  No GL driver here can render
    that (llvmpipe's limit is 8192), so the oracle restatement is the checker.  Then the same panorama delivered to a
    host buffer by 8 azimuth wedges one after the other (what 8 ranks do concurrently, tests/test_multigpu.py) must
    equal the whole render bit for bit."""
    import horizonator_b200 as hz
    from horizonator_b200 import sharding
    from oracle.binding import Oracle
    W, H = 36000, 4000
    az0, az1 = -180.0 + 180.0 / W, 180.0 - 180.0 / W
    h = hz.horizonator(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles_c2, render_radius_m=150000.)
    h.set_zextents(100., 150000.)
    h.pan_zoom(az0, az1)
    img = hz.pinned_array((H, W, 3), np.uint8)
    rng = hz.pinned_array((H, W), np.float32)
    h.render_into(img, rng)
    o = Oracle(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles_c2, render_radius_m=150000., threads=os.cpu_count() or 1)
    img_o, rng_o = o.render(az0, az1, znear=100., zfar=150000.)
    del o
    s = _compare_in_column_chunks(img, rng, img_o, rng_o)
    print("config 4 at spec vs oracle:", s)
    assert s["hit_fraction_ref"] > 0.02
    del img_o, rng_o
    # 8 wedges, sequentially, into one host buffer (pageable here: the strided 2-D copies must cope with that too)
    wi = np.zeros((H, W, 3), np.uint8)
    wr = np.zeros((H, W), np.float32)
    edges = [((W * g // 8) // 4) * 4 for g in range(8)] + [W]
    for g in range(8):
        h.render_wedge_host(edges[g], edges[g + 1], wi, wr)
    assert np.array_equal(wi, img) and np.array_equal(wr, rng)
    # the shared-buffer driver with a world of one
    hp = sharding.HostPanorama(h)
    pi, pr = hp.render()
    assert np.array_equal(pi, img) and np.array_equal(pr, rng)
    hp.close()


def test_widest_panorama_the_reference_can_render_matches_llvmpipe(tiles_c1):
    """8192 columns: the renderbuffer limit of the image's llvmpipe (horizonator-lib.c:633 fails with GL_INVALID_VALUE
    beyond it), i.e. the widest panorama the reference itself can produce here; recorded from the reference on llvmpipe
    by tests/golden/make_golden_llvmpipe_wide.py (terrain/sky of every pixel; range and red of every 8th column)."""
    import horizonator_b200 as hz
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wide8192_llvmpipe.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/wide8192_llvmpipe.npz was not generated")
    g = np.load(path)
    W, H, R, az0, az1, zn, zf, step = (float(x) for x in g["params"])
    W, H, R, step = int(W), int(H), int(R), int(step)
    C1_LAT, C1_LON = 35.0 + 1.0 / 2400.0, -117.0 + 1.0 / 2400.0
    h = hz.horizonator(C1_LAT, C1_LON, W, H, dir_dems=tiles_c1, render_radius_cells=R)
    img, rng = h.render(az0, az1, znear=zn, zfar=zf)
    hit_ref = np.unpackbits(g["hit_bits"], axis=1)[:, :W].astype(bool)
    hit = rng > 0
    agree = float((hit == hit_ref).mean())
    print("8192-wide vs llvmpipe: coverage agreement %.6f, %d pixels differ, terrain %.4f" % (agree, int((hit != hit_ref).sum()), hit_ref.mean()))
    assert agree >= 0.995
    # range / red on the recorded columns, with the silhouette rule of compare_renders
    img_sub = np.zeros((H, g["ranges_sub"].shape[1], 3), np.uint8)
    img_sub[..., 2] = g["red_sub"]; img_sub[..., 0] = np.where(g["ranges_sub"] > 0, 0, 255)
    # (neighbouring recorded columns are 8 pixels apart: treat every column on its own for the silhouette test by
    # comparing vertically only -- compare_renders' 3x3 neighbourhood across the gaps just marks more silhouettes)
    s = compare_renders(np.ascontiguousarray(img[:, ::step]), np.ascontiguousarray(rng[:, ::step]), img_sub, g["ranges_sub"])
    print("8192-wide vs llvmpipe, every %dth column:" % step, s)
    assert s["coverage_agreement"] >= 0.995 and s["red_max_diff_where_agree"] <= 1 and s["sky_ok"]
    both = (rng[:, ::step] > 0) & (g["ranges_sub"] > 0)
    rel = np.abs(rng[:, ::step][both].astype(np.float64) - g["ranges_sub"][both]) / g["ranges_sub"][both]
    assert float((rel > 1e-4).mean()) < 0.005, float((rel > 1e-4).mean())


@pytest.mark.parametrize("W,H", [(3600, 600), (1234, 77)])
def test_pageable_destinations_get_what_page_locked_ones_get(tiles_c2, W, H):
    """horizonator_render_offscreen() into ordinary pageable arrays -- what the reference's Python binding passes --
    goes through the chunked staging pipeline (several host threads); into page-locked arrays by plain DMA.  Same
    bytes either way, also when only one of the two outputs is asked for."""
    import horizonator_b200 as hz
    h = hz.horizonator(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles_c2, render_radius_cells=600)
    h.set_zextents(100., 60000.)
    h.pan_zoom(-180.05, 179.95)
    pi, pr = hz.pinned_array((H, W, 3), np.uint8), hz.pinned_array((H, W), np.float32)
    h.render_into(pi, pr)
    assert (pr > 0).mean() > 0.01
    for rep in range(3):
        ai, ar = np.full((H, W, 3), 7, np.uint8), np.full((H, W), 7, np.float32)
        h.render_into(ai, ar)
        assert np.array_equal(ai, pi) and np.array_equal(ar, pr), rep
    ai = np.full((H, W, 3), 7, np.uint8)
    h.render_into(ai, None)
    assert np.array_equal(ai, pi)
    ar = np.full((H, W), 7, np.float32)
    h.render_into(None, ar)
    assert np.array_equal(ar, pr)


def test_lod_is_opt_in_and_stays_within_subpixel_error(tiles_c2):
    """Opt-in level of detail (horizonator_set_lod; N4): off by default -- every other test runs without it -- and
    switching it off again restores the full render bit for bit.  On (coarser cells up to half a pixel across), the
    image may differ from the full render only by sub-pixel shifts of far silhouettes: reported here, and bounded."""
    import torch
    import horizonator_b200 as hz
    W, H = 3600, 600
    h = hz.horizonator(C2_LAT, C2_LON, W, H, SRTM1=True, dir_dems=tiles_c2, render_radius_m=150000.)
    h.set_zextents(100., 150000.)
    stream = torch.cuda.Stream()
    d_img = torch.empty((2, H, W, 3), dtype=torch.uint8, device="cuda")
    d_rng = torch.empty((2, H, W), dtype=torch.float32, device="cuda")
    views = {"c2": (C2_LAT, C2_LON, -180.05, 179.95), "eye_12km": (C2_LAT, C2_LON, -180.05, 179.95, 12000.),
             "grid": (33.5 + 1.5 / 8 + 1.0 / 7200.0, -117.5 + 5.5 / 8 + 1.0 / 7200.0, -180.05, 179.95)}

    def render(v, slot):
        t = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            h.render_batch_device([v], d_img[slot].data_ptr(), d_rng[slot].data_ptr(), stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1))
        return d_img[slot].cpu().numpy(), d_rng[slot].cpu().numpy(), min(t)

    for name, v in views.items():
        h.set_lod(0.)
        fi, fr, t_full = render(v, 0)
        for px in (0.5, 1.0, 2.0):
            h.set_lod(px)
            li, lr, t_lod = render(v, 1)
            s = compare_renders(li, lr, fi, fr)
            both = (lr > 0) & (fr > 0)
            rel = np.abs(lr[both].astype(np.float64) - fr[both]) / fr[both]
            print("LOD %.1f px vs full, %s: %.3f ms -> %.3f ms; coverage agreement %.5f, agreement incl. range %.5f, "
                  "off-silhouette %d, range error median %.2e / p99 %.2e" %
                  (px, name, t_full, t_lod, s["coverage_agreement"], s["agreement"], s["off_silhouette"], float(np.median(rel)),
                   float(np.quantile(rel, 0.99))))
            if px == 0.5:
                assert s["coverage_agreement"] >= 0.995 and s["sky_ok"] and s["bg_channels_ok"]
                assert float(np.quantile(rel, 0.99)) < 5e-3          # a few metres at a few kilometres
        h.set_lod(0.)
        bi, br, _ = render(v, 1)
        assert np.array_equal(bi, fi) and np.array_equal(br, fr)
    # batches take the same path -- also when the views' windows differ (the level changes at rings that depend on the
    # window's width: the first view's band limits must not shape the launches of the others)
    h.set_lod(0.5)
    bviews = [(C2_LAT, C2_LON, 40., 41.), views["c2"], views["grid"], (C2_LAT, C2_LON, -60., 120.), (C2_LAT + 0.02, C2_LON, 0., 20.)]
    for rep in range(2):
        bi, br = h.render_batch(bviews)
        for k, v in enumerate(bviews):
            h.move(v[0], v[1])
            h.pan_zoom(v[2], v[3])
            wi = hz.pinned_array((H, W, 3), np.uint8); wr = hz.pinned_array((H, W), np.float32)
            h.render_into(wi, wr)
            assert np.array_equal(bi[k], wi) and np.array_equal(br[k], wr), (rep, k)
