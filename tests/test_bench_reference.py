"""bench.py --impl reference (the reference's CPU path through oracle/_ref, else the port) runs without a GPU: the JSON line
carries the keys the driver reads, on the workload's own metric / unit / config (here on a small override grid)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "32", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("leaf cells/sec") and d["unit"] == "cells/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["config"]["grid"] == [32, 32, 32] and d["config"]["cells"] == 32 ** 3 and "workload" in d["config"]
    assert d["value"] > 0 and abs(d["value"] - 32 ** 3 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
