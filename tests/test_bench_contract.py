"""CPU: the parts of bench.py's contract that need no GPU -- the reference arm (`--impl reference`: the CPU oracle on the
bounded sample of the bench workload) prints ONE JSON line with the keys the driver reads, and marks its number as what
it is (a port, extrapolated from the sample)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"] == "Sim3 LM iterations/s" and d["unit"] == "LM iterations/s" and d["dtype"] == "f64"
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["config"]["workload"].startswith("s1m")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["extrapolated"] is True and cb["cores"] >= 1 and "sample" in cb
    assert abs(cb["value"] - d["value"]) <= 1e-12 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and e["unit"] == d["unit"]
    assert abs(e["value"] - d["value"]) <= 1e-12 * d["value"]


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun (N > 1) rank 0 alone runs the CPU arm; the other ranks exit 0 without output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
