"""The GL-free command-line renderer (cli/horizonator-standalone.c) on the GPU: its PNG and range
dump equal what the library returns for the same view under the reference tool's conventions
(standalone.c:403-411: pixel-centre azimuths, 20-degree default field of view), and its annotation geometry
follows annotator.c:280-348."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import C1_LAT, C1_LON

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "horizonator_b200", "bin", "horizonator-standalone")


def test_cli_png_and_ranges_match_library(tiles_c1, tmp_path):
    from PIL import Image
    import horizonator_b200 as hz
    W = 900
    png, f32 = str(tmp_path / "out.png"), str(tmp_path / "out.f32")
    az_c, az_r, zfar = 60.0, 50.0, 30000.0
    r = subprocess.run([CLI, "--width", str(W), "--image", png, "--ranges", f32, "--zfar", str(zfar), "--dirdems", tiles_c1,
                        "%.9f" % C1_LAT, "%.9f" % C1_LON, str(az_c), str(az_r)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the tool's conventions, in float like the tool
    az_r_f = np.float32(az_r)
    az_r_f = az_r_f + np.float32(2.0 * float(az_r_f) / float(np.float32(W - 1))) / np.float32(2.0)
    H = int(np.rint(np.float32(W) * np.float32(20.0) / az_r_f))
    rgb = np.asarray(Image.open(png))
    assert rgb.shape == (H, W, 3)
    rng = np.fromfile(f32, dtype=np.float32).reshape(H, W)

    # the same view through the C API as the tool drives it: radius = zfar metres, automatic eye height
    ctx = hz.context_t()
    z = C.c_float(-1.0)
    assert hz.lib.horizonator_init(C.byref(ctx), np.float32(C1_LAT), np.float32(C1_LON), C.byref(z), W, H, -1, zfar, True, False,
                                   False, os.fsencode(tiles_c1), None, None, None, False)
    try:
        assert hz.lib.horizonator_set_zextents(C.byref(ctx), 100., zfar, 100., zfar)
        assert hz.lib.horizonator_pan_zoom(C.byref(ctx), np.float32(az_c) - az_r_f, np.float32(az_c) + az_r_f)
        img = np.empty((H, W, 3), np.uint8)
        want = np.empty((H, W), np.float32)
        assert hz.lib.horizonator_render_offscreen(C.byref(ctx), img.ctypes.data, want.ctypes.data)
    finally:
        hz.lib.horizonator_deinit(C.byref(ctx))
    assert (want > 0).mean() > 0.02
    assert np.array_equal(rng, want)
    assert np.array_equal(rgb, img[..., ::-1])            # PNG is RGB, the library returns B,G,R


def test_cli_labels_follow_annotator_rules(tiles_c1, tmp_path):
    import horizonator_b200 as hz
    W, H = 1200, 300
    png, f32, labels = str(tmp_path / "o.png"), str(tmp_path / "o.f32"), str(tmp_path / "labels.json")
    pois = tmp_path / "pois.csv"
    # a first run without POIs to learn where visible terrain is
    base = [CLI, "--width", str(W), "--height", str(H), "--zfar", "40000", "--dirdems", tiles_c1]
    where = ["%.9f" % C1_LAT, "%.9f" % C1_LON, "0", "90"]
    assert subprocess.run(base + ["--image", png, "--ranges", f32] + where, capture_output=True).returncode == 0
    rng = np.fromfile(f32, dtype=np.float32).reshape(H, W)
    az_r = np.float32(90.0) + np.float32(2.0 * 90.0 / (W - 1)) / np.float32(2.0)
    az0, az1 = float(np.float32(0.0) - az_r), float(np.float32(0.0) + az_r)
    # visible POIs: skyline points (the topmost terrain pixel of a column, like the summits the reference annotates --
    # its search walks down from 6 rows above the predicted position and gives up as soon as the range error grows,
    # annotator.c:314-347, which suits points with sky above them), between 2 and 30 km
    top = np.argmax(rng > 0, axis=0)
    cols = [c for c in range(5, W - 5) if rng[top[c], c] > 2000. and rng[top[c], c] < 30000. and top[c] > 8]
    pick_cols = np.random.default_rng(0).choice(cols, 12, replace=False)
    ys = np.array([top[c] for c in pick_cols]); xs = np.array(pick_cols)
    pick = range(12)
    lines, expect = [], []
    la, lo = C.c_float(), C.c_float()
    h = hz.horizonator(C1_LAT, C1_LON, W, H, dir_dems=tiles_c1, render_radius_m=40000.)
    eye_z = h.move(C1_LAT, C1_LON)
    for k, p in enumerate(pick):
        x, y = int(xs[p]), int(ys[p])
        # reported range = slant / cos(el) (reference quirk Q1), so the horizontal distance is range * cos(el)^2
        assert hz.lib.horizonator_unproject(C.byref(la), C.byref(lo), x, y, -1.,
                                            float(rng[y, x]) * np.cos(_el(y, H, W, az0, az1)) ** 2, float(np.float32(C1_LAT)),
                                            np.cos(np.radians(float(np.float32(C1_LAT)))), float(np.float32(C1_LON)),
                                            az0, az1, W, H)
        # the height of the rendered surface point itself: eye height + horizontal distance * tan(elevation)
        ele = eye_z + float(rng[y, x]) * np.cos(_el(y, H, W, az0, az1)) ** 2 * np.tan(_el(y, H, W, az0, az1))
        lines.append("seen%d,%.7f,%.7f,%.2f" % (k, la.value, lo.value, ele))
        expect.append(("seen%d" % k, x, y))
    # hidden POIs: far below the terrain, behind the viewer's window, too close, too far
    lines.append("buried,%.7f,%.7f,-5000" % (C1_LAT + 0.1, C1_LON))
    lines.append("behind,%.7f,%.7f,1000" % (C1_LAT - 0.1, C1_LON))
    lines.append("tooclose,%.7f,%.7f,%d" % (C1_LAT + 0.001, C1_LON, int(eye_z)))
    lines.append("toofar,%.7f,%.7f,3000" % (C1_LAT + 1.5, C1_LON))
    pois.write_text("# name,lat,lon,ele\n" + "\n".join(lines) + "\n")
    r = subprocess.run(base + ["--image", png, "--pois", str(pois), "--labels", labels] + where, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = {d["name"]: d for d in json.load(open(labels))}
    assert not ({"buried", "behind", "tooclose", "toofar"} & set(got))
    found = 0
    for name, x, y in expect:
        if name in got:
            found += 1
            assert abs(got[name]["x"] - x) <= 1.0 and abs(got[name]["y"] - y) <= 2.0, (name, got[name], x, y)
            assert abs(got[name]["range_m"] - got[name]["range_rendered_m"]) < 500.0
    print("labels found", found, "of", len(expect), sorted(got))
    assert found >= 11, (found, sorted(got), open(labels).read(), lines)


def _el(row, H, W, az0, az1):
    """elevation of an image row"""
    return np.radians((1 - (2 * row + 1) / H) * (az1 - az0) / (2 * (W / H)))
