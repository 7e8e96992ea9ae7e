"""bench.py's reference arm on the CPU: exactly one JSON line on stdout with the keys the driver reads.  Runs the
arm on the GL restatement with a one-second budget (one full C2 panorama); the llvmpipe arm is the same code path
with a slower renderer and is exercised by tests/test_llvmpipe.py."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, HZ_REF_GL="restated", HZ_REF_BUDGET_S="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "2", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "panoramas/sec (SRTM1, 3600x600 px)" and d["unit"] == "panoramas/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - 1000.0 / d["ms_per_step"]) < 1e-9 * d["value"] + 1e-12
    assert d["steps"] >= 1 and "workload" in d["config"] and "configs[1]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"]
    assert "sample" in cb and "C2 panorama" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""
