"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port of the reference's CPU eager path) prints ONE JSON line
with the contract's keys and the same metric / workload strings as the product arm."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, REPO)
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "tokens/s"
    assert d["config"]["workload"] == bench.WORKLOAD and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit="tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
