"""bench.py's reference arm runs on host cores only (the unmodified reference's CPU backend out of oracle/_ref), so
its JSON contract can be checked without a GPU: one line, the keys the driver reads, `impl: reference`, an `e2e`
object that repeats the line's own value with zero transfer bytes, and a `cpu_baseline` describing the run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libgtref.so")),
    reason="oracle/_ref/libgtref.so not built")
def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
        capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["metric"].startswith("Mpts/s vert_adv 256x256x80") and d["unit"] == "Mpts/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] * 1e6 - 256 * 256 * 80) < 1e-3 * 256 * 256 * 80
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == d["value"] and cb["cores"] >= 1 and "sample" in cb
    assert d["vs_baseline"] is None and d["dtype"] == "f64"
