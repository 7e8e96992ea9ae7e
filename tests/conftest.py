import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_build_module():
    """horizonator_b200/build.py loaded by path: importing the package itself needs an up-to-date library."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("hz_build", os.path.join(ROOT, "horizonator_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build what is missing: the product library (nvcc), the synthetic-tile generator and the oracle (gcc)."""
    hb = load_build_module()
    hb.build_library()
    hb.build_synth()
    hb.build_cli()
    from oracle import binding
    binding.build(ref=os.path.isdir("/root/reference"))


@pytest.fixture(scope="session")
def tiles_c1(tmp_path_factory):
    """BASELINE config 1 terrain: 2x2 synthetic SRTM3 tiles N34..N35 x W118..W117 (seed 7, with voids)."""
    from tools import synth
    d = os.environ.get("HZ_TILES_C1") or str(tmp_path_factory.mktemp("dems_srtm3"))
    return synth.config1_tiles(d)


@pytest.fixture(scope="session")
def tiles_holes(tmp_path_factory):
    """Same block with N35W118 missing and N34W117 present but zero-length (both read as elevation 0)."""
    from tools import synth
    d = str(tmp_path_factory.mktemp("dems_holes"))
    synth.write_tiles(d, (34, 35), (-118, -117), seed=7, skip=((35, -118), (34, -117)))
    open(os.path.join(d, synth.tile_name(34, -117)), "wb").close()
    return d


# viewer of BASELINE config 1: block centre + half a cell
C1_LAT = 35.0 + 1.0 / 2400.0
C1_LON = -117.0 + 1.0 / 2400.0
