"""The accuracy contract of csrc/hz_math.cuh, measured on the GPU against double precision:

    hz_atan2_az(e, n)  : <= 2.5 ulp of the result over all four quadrants
    hz_atan_el(h, d2)  : <= 3.5 ulp of the result (incl. the steep branch |h| > d and the d2 -> 0 cases;
                         measured on B200: 3.03 ulp at worst, 0.44 ulp on average)

The functions replace GLSL's atan() in vertex.glsl:136,153; the oracle uses glibc's atan2f (<= 1 ulp).  A few ulp of
an angle are ~1e-7 relative: four orders of magnitude inside the 1/256-pixel snapping of the rasteriser for any image
up to 36000 columns, which is what the parity bar needs.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ulp_error(got, want64):
    """|got - want| in units of the float32 spacing at |want|."""
    want32 = want64.astype(np.float32)
    ulp = np.spacing(np.abs(want32)).astype(np.float64)
    ulp = np.maximum(ulp, np.float64(np.finfo(np.float32).tiny))
    return np.abs(got.astype(np.float64) - want64) / ulp


def _probe(hz, e, n, h, d2):
    e, n, h, d2 = (np.ascontiguousarray(a, np.float32) for a in (e, n, h, d2))
    az, el = np.empty_like(e), np.empty_like(e)
    assert hz.lib.horizonator_debug_device_math(len(e), e.ctypes.data, n.ctypes.data, h.ctypes.data, d2.ctypes.data,
                                                az.ctypes.data, el.ctypes.data)
    return az, el


@pytest.fixture(scope="module")
def hz():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import horizonator_b200
    return horizonator_b200


def test_azimuth_atan2_within_2p5_ulp_over_all_quadrants(hz):
    rs = np.random.default_rng(1)
    n = 2_000_000
    # metres east/north of the eye as the renderer sees them: 1 m .. 300 km, both signs, plus ratios near 0, 1 and inf
    mag = lambda: np.exp(rs.uniform(np.log(1.0), np.log(3e5), n)) * rs.choice([-1.0, 1.0], n)
    e, nn = mag(), mag()
    k = n // 10
    e[:k] = nn[:k] * (1.0 + rs.uniform(-1e-3, 1e-3, k))            # near the 45-degree fold of the polynomial
    e[k:2 * k] = nn[k:2 * k] * rs.uniform(-1e-6, 1e-6, k)           # nearly due north / south
    nn[2 * k:3 * k] = e[2 * k:3 * k] * rs.uniform(-1e-6, 1e-6, k)   # nearly due east / west
    one = np.ones(n, np.float32)
    az, _ = _probe(hz, e, nn, one, one)
    e32, n32 = e.astype(np.float32).astype(np.float64), nn.astype(np.float32).astype(np.float64)
    err = _ulp_error(az, np.arctan2(e32, n32))
    print("hz_atan2_az: max %.3f ulp, mean %.3f ulp over %d arguments" % (err.max(), err.mean(), n))
    assert err.max() <= 2.5
    # exact values on the axes and the sign convention (0 = north, +pi/2 = east, +-pi = south)
    az, _ = _probe(hz, [0., 1., 0., -1., 0., -0.0], [1., 0., -1., 0., 0., -1.], [1.] * 6, [1.] * 6)
    assert az[0] == 0.0 and az[1] == np.float32(np.pi / 2) and az[2] == np.float32(np.pi) and az[3] == -np.float32(np.pi / 2)
    assert az[4] == 0.0 and az[5] == -np.float32(np.pi)


def test_elevation_atan_within_3p5_ulp_incl_steep_and_degenerate(hz):
    rs = np.random.default_rng(2)
    n = 2_000_000
    d = np.exp(rs.uniform(np.log(0.5), np.log(3e5), n))
    h = np.exp(rs.uniform(np.log(1e-3), np.log(2e4), n)) * rs.choice([-1.0, 1.0], n)
    k = n // 10
    h[:k] = d[:k] * rs.uniform(0.9, 1.1, k) * rs.choice([-1.0, 1.0], k)        # around 45 degrees: the branch point
    h[k:2 * k] = d[k:2 * k] * np.exp(rs.uniform(0, np.log(1e4), k))            # steeper: right under / above the eye
    d2 = (d.astype(np.float32).astype(np.float64)) ** 2
    d2_32 = d2.astype(np.float32)
    one = np.ones(n, np.float32)
    _, el = _probe(hz, one, one, h, d2_32)
    want = np.arctan2(h.astype(np.float32).astype(np.float64), np.sqrt(d2_32.astype(np.float64)))
    err = _ulp_error(el, want)
    print("hz_atan_el: max %.3f ulp, mean %.3f ulp over %d arguments" % (err.max(), err.mean(), n))
    assert err.max() <= 3.5
    # d2 == 0: straight up / down, and atan(0, 0) = 0
    _, el = _probe(hz, [1.] * 4, [1.] * 4, [5., -5., 0., -0.0], [0., 0., 0., 0.])
    assert el[0] == np.float32(np.pi / 2) and el[1] == -np.float32(np.pi / 2) and el[2] == 0.0 and el[3] == 0.0
