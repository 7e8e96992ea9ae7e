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
