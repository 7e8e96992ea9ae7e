"""Command-line front end without a GPU: usage errors and the loud failure when no CUDA device exists
(cli/horizonator-standalone.c; rendering itself is covered by tests/test_gpu_cli.py)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "horizonator_b200", "bin", "horizonator-standalone")


def _run(*args):
    return subprocess.run([CLI] + list(args), capture_output=True, text=True, timeout=120)


def test_usage_errors():
    p = _run()
    assert p.returncode != 0 and "Need exactly 4 non-option arguments" in p.stderr + p.stdout
    p = _run("--bogus", "35", "-117", "0", "10")
    assert p.returncode != 0
    p = _run("--image", "/tmp/never_written.png", "35", "-117", "0", "10")      # --width is mandatory
    assert p.returncode != 0 and not os.path.exists("/tmp/never_written.png")


def test_fails_loudly_without_a_cuda_device(tiles_c1, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = tmp_path / "x.png"
    p = _run("--width", "360", "--image", str(out), "--dirdems", tiles_c1, "35.0", "-117.0", "0", "10")
    assert p.returncode != 0 and not out.exists()
    assert "No usable CUDA device" in p.stderr and "no CPU path" in p.stderr
