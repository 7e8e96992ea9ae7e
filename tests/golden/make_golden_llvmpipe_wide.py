#!/usr/bin/env python
"""Generates tests/golden/wide8192_llvmpipe.npz: the widest panorama the REFERENCE can render at all (llvmpipe's
renderbuffer limit is 8192 -- GL_INVALID_VALUE beyond, lib:633; horizonator-lib.c:617-666 creates the FBO at the panorama's size), by the reference
itself on Mesa llvmpipe -- the pin for the wedge-sharded giant panoramas of BASELINE configs[3], whose own size
(36000 x 4000) no GL driver here can do.

    python tests/golden/make_golden_llvmpipe_wide.py        (development container: needs /root/reference)

Scene: BASELINE configs[0]'s DEM (2x2 synthetic SRTM3 tiles, R = 1200 cells), full circle, 8192 x 910 (the 9:1
aspect of 36000 x 4000).  To keep the fixture small the file keeps: the terrain/sky bitmap of
EVERY pixel (packed bits), and range + red channel of every 8th column.
"""
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
W, H, R = 8192, 910, 1200
AZ0, AZ1, ZNEAR, ZFAR = -180.0 + 180.0 / W, 180.0 - 180.0 / W, 100., 100000.
STEP = 8


def main():
    if not os.path.isdir("/root/reference"):
        raise SystemExit("needs /root/reference (run in the development container)")
    binding.build(ref=True)
    tiles = synth.config1_tiles(os.path.join(tempfile.mkdtemp(prefix="hz_golden_"), "c1"))
    r = binding.MesaReference(C1_LAT, C1_LON, W, H, dir_dems=tiles, render_radius_cells=R, threads=os.cpu_count() or 1)
    img, rng = r.render(AZ0, AZ1, znear=ZNEAR, zfar=ZFAR)
    version, renderer = r.gl_strings()
    assert "llvmpipe" in renderer, renderer
    hit = rng > 0
    assert 0.01 < hit.mean() < 0.9
    # everything but the red channel is a function of hit/sky (fragment.glsl:16, lib:185): checked here, not stored
    assert np.all(img[..., 1] == 0) and np.all(img[..., 0][hit] == 0) and np.all(img[..., 0][~hit] == 255)
    np.savez_compressed(os.path.join(HERE, "wide8192_llvmpipe.npz"),
                        hit_bits=np.packbits(hit, axis=1), ranges_sub=rng[:, ::STEP].copy(), red_sub=img[:, ::STEP, 2].copy(),
                        params=np.array([W, H, R, AZ0, AZ1, ZNEAR, ZFAR, STEP], np.float64), viewer_z=np.float32(r.viewer_z))
    print("written: %dx%d, terrain fraction %.4f, %s / %s" % (W, H, hit.mean(), version, renderer))
    r.close()


if __name__ == "__main__":
    main()
