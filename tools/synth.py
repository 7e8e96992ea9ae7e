"""Seeded synthetic SRTM tiles (tests and benchmarks run without network or real DEMs).

Thin ctypes wrapper over tools/synth_hgt.c (built into horizonator_b200/lib/libsynth.so by
horizonator_b200.build.build_synth()).
"""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "horizonator_b200", "lib", "libsynth.so")
        if not os.path.exists(path):
            import sys
            sys.path.insert(0, ROOT)
            from horizonator_b200.build import build_synth
            build_synth()
        _LIB = C.CDLL(path)
        _LIB.synth_write_tile.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int]
        _LIB.synth_write_tile.restype = C.c_int
    return _LIB


def tile_name(lat, lon):
    return "%s%02d%s%03d.hgt" % ("N" if lat >= 0 else "S", abs(lat), "E" if lon >= 0 else "W", abs(lon))


def write_tiles(directory, lats, lons, srtm1=False, seed=7, voids=True, skip=()):
    """Writes one .hgt per (lat, lon) SW corner; `skip` lists (lat, lon) pairs to leave out."""
    os.makedirs(directory, exist_ok=True)
    cpd = 3600 if srtm1 else 1200
    for lat in lats:
        for lon in lons:
            if (lat, lon) in skip:
                continue
            path = os.path.join(directory, tile_name(lat, lon))
            if os.path.exists(path) and os.path.getsize(path) == (cpd + 1) ** 2 * 2:
                continue
            rc = _lib().synth_write_tile(os.fsencode(directory), lat, lon, cpd, seed, 1 if voids else 0)
            if rc != 0:
                raise RuntimeError("synth_write_tile(%d,%d) failed: %d" % (lat, lon, rc))
    return directory


# BASELINE.json configs
def config1_tiles(directory, seed=7):
    """C1: 2x2 SRTM3 tiles N34..N35 x W118..W117; viewer (35+1/2400, -117+1/2400), R=1200."""
    return write_tiles(directory, (34, 35), (-118, -117), srtm1=False, seed=seed)


def config2_tiles(directory, seed=7):
    """C2: 4x4 SRTM1 tiles N32..N35 x W119..W116; viewer (34+1/7200, -117+1/7200), 150 km."""
    return write_tiles(directory, (32, 33, 34, 35), (-119, -118, -117, -116), srtm1=True, seed=seed)
