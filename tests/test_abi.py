"""The drop-in boundary, checked without a GPU: struct layout of include/*.h (against the numbers measured on
the reference's own headers, SURVEY.md 8b, and against the reference's headers themselves where
/root/reference exists), the ctypes mirror, and that libhorizonator.so loads and exports every symbol the
headers declare.  No compute call is made here."""
import ctypes as C
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"

PROBE = r"""
#include <stddef.h>
#include <stdio.h>
#include "horizonator.h"
#define O(f) printf(#f " %zu\n", offsetof(horizonator_context_t, f))
#define D(f) printf("dems." #f " %zu\n", offsetof(horizonator_dem_context_t, f))
int main(void)
{
    printf("sizeof_context %zu\n", sizeof(horizonator_context_t));
    printf("sizeof_dems %zu\n", sizeof(horizonator_dem_context_t));
    O(Ntriangles); O(render_texture); O(use_glut); O(glut_window); O(uniform_aspect); O(uniform_zfar_color);
    O(program); O(viewer_lat); O(viewer_lon); O(dems);
    O(offscreen.inited); O(offscreen.frameBufID); O(offscreen.renderBufID); O(offscreen.depthBufID);
    O(offscreen.width); O(offscreen.height);
    D(dems); D(mmap_sizes); D(mmap_fd); D(origin_dem_lon_lat); D(origin_dem_cellij); D(Ndems_ij);
    D(radius_cells); D(cells_per_deg);
    printf("max_Ndems_ij %d\n", (int)max_Ndems_ij);
    printf("znear_default %g\nzfar_default %g\n", (double)HORIZONATOR_ZNEAR_DEFAULT, (double)HORIZONATOR_ZFAR_DEFAULT);
    return 0;
}
"""

# gcc 13.3 x86-64 on the reference's horizonator.h / dem.h
EXPECTED = {
    "sizeof_context": 472, "sizeof_dems": 352,
    "Ntriangles": 0, "render_texture": 4, "use_glut": 5, "glut_window": 8, "uniform_aspect": 12,
    "uniform_zfar_color": 76, "program": 80, "viewer_lat": 84, "viewer_lon": 88, "dems": 96,
    "offscreen.inited": 448, "offscreen.frameBufID": 452, "offscreen.renderBufID": 456, "offscreen.depthBufID": 460,
    "offscreen.width": 464, "offscreen.height": 468,
    "dems.dems": 0, "dems.mmap_sizes": 128, "dems.mmap_fd": 256, "dems.origin_dem_lon_lat": 320,
    "dems.origin_dem_cellij": 328, "dems.Ndems_ij": 336, "dems.radius_cells": 344, "dems.cells_per_deg": 348,
    "max_Ndems_ij": 4, "znear_default": 100, "zfar_default": 40000,
}


def _probe(include_dir, tmp_path, tag):
    src = tmp_path / ("probe_%s.c" % tag)
    exe = tmp_path / ("probe_%s" % tag)
    src.write_text(PROBE)
    subprocess.run([GCC, "-std=gnu99", "-I", include_dir, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    return {k: float(v) for k, v in (line.split() for line in out.splitlines())}


def test_struct_layout_matches_survey(tmp_path):
    got = _probe(INCLUDE, tmp_path, "ours")
    assert got == {k: float(v) for k, v in EXPECTED.items()}


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference headers not present on this machine")
def test_struct_layout_matches_reference_headers(tmp_path):
    assert _probe(INCLUDE, tmp_path, "ours") == _probe("/root/reference", tmp_path, "ref")


def test_ctypes_mirror_matches_headers():
    import horizonator_b200 as hz
    assert C.sizeof(hz.context_t) == EXPECTED["sizeof_context"]
    assert C.sizeof(hz.dem_context_t) == EXPECTED["sizeof_dems"]
    for name in ("Ntriangles", "render_texture", "use_glut", "glut_window", "program", "viewer_lat", "viewer_lon", "dems"):
        assert getattr(hz.context_t, name).offset == EXPECTED[name], name
    assert hz.context_t.uniforms.offset == EXPECTED["uniform_aspect"]
    assert hz.context_t.offscreen.offset == EXPECTED["offscreen.inited"]
    for name in ("mmap_sizes", "mmap_fd", "origin_dem_lon_lat", "origin_dem_cellij", "Ndems_ij", "radius_cells",
                 "cells_per_deg"):
        assert getattr(hz.dem_context_t, name).offset == EXPECTED["dems." + name], name
    assert C.sizeof(hz.view_t) == 20


def _declared_functions():
    names = set()
    for h in glob.glob(os.path.join(INCLUDE, "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        names |= set(re.findall(r"\b(horizonator_[a-z0-9_]+)\s*\(", text))
    names.discard("horizonator_context_isvalid")      # static inline in horizonator.h
    return names


def test_library_exports_every_declared_symbol():
    import horizonator_b200 as hz
    declared = _declared_functions()
    assert len(declared) >= 26
    raw = C.CDLL(hz.LIBRARY_PATH)
    missing = [n for n in sorted(declared) if not hasattr(raw, n)]
    assert not missing, missing
    assert set(hz.EXPORTED_SYMBOLS) == declared
    # dynamic symbol table: C linkage (no mangled product entry points), SONAME as the reference's ABI 0
    dyn = subprocess.run(["nm", "-D", "--defined-only", hz.LIBRARY_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in dyn.splitlines() if " T " in line}
    assert declared <= exported
    soname = subprocess.run(["readelf", "-d", hz.LIBRARY_PATH], capture_output=True, text=True, check=True).stdout
    assert "libhorizonator.so.0" in soname


def test_reference_callers_need_only_exported_symbols():
    """The six entry points the reference's Python binding imports (verified with nm there)."""
    import horizonator_b200 as hz
    for n in ("horizonator_init", "horizonator_deinit", "horizonator_move", "horizonator_pan_zoom",
              "horizonator_set_zextents", "horizonator_render_offscreen"):
        assert n in hz.EXPORTED_SYMBOLS
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "pywrap", "horizonator*.so"))
    if so:
        und = subprocess.run(["nm", "-D", "--undefined-only", so[0]], capture_output=True, text=True, check=True).stdout
        needed = {l.split()[-1] for l in und.splitlines() if "horizonator_" in l}
        assert needed and needed <= set(hz.EXPORTED_SYMBOLS), needed


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: no product source may import, link or execute it."""
    pkg = os.path.join(ROOT, "horizonator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
    import horizonator_b200 as hz
    needed = subprocess.run(["readelf", "-d", hz.LIBRARY_PATH], capture_output=True, text=True, check=True).stdout
    assert "oracle" not in needed


def test_no_cpu_fallback_without_a_device(tmp_path):
    """Without a CUDA device the constructor must fail loudly (no CPU rendering path exists)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import horizonator_b200 as hz
    with pytest.raises(RuntimeError):
        hz.horizonator(35.0004, -116.9996, 64, 16, dir_dems=str(tmp_path), render_radius_cells=8)
    ctx = hz.context_t()
    assert not hz.lib.horizonator_init(C.byref(ctx), 35.0004, -116.9996, None, 64, 16, 8, -1.0, True, False, False,
                                       os.fsencode(str(tmp_path)), None, None, None, False)
    assert ctx.Ntriangles == 0                               # horizonator_context_isvalid() == false
    hz.lib.horizonator_deinit(C.byref(ctx))                  # safe on a failed/zeroed context (pywrap.c:127-131)
    assert not hz.lib.horizonator_render_offscreen(C.byref(ctx), None, None)
    assert not hz.lib.horizonator_pan_zoom(C.byref(ctx), 0., 90.)
