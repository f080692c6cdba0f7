"""bench.py's output contract, checked without a GPU: the committed line of the last measured run carries every key
the driver reads, and the reference arm (CPU oracle port) prints a well-formed line here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"]


def test_committed_bench_line_has_every_contract_key():
    with open(os.path.join(ROOT, "profiles", "r2_bench_line_bigvgan.json")) as f:
        d = json.loads(f.read())
    for k in BASE_KEYS + ["clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"]:
        assert k in d, k
    assert d["metric"] == "audio-seconds/sec" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None          # BASELINE.md publishes no number for this metric
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] != d["value"]   # the end-to-end leg is measured, not copied
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("port", "reference")
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    assert r["traffic"] is None or abs(r["traffic"] - t["bigvgan_b64_f500_f16"]) < 0.01 * r["traffic"]


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--frames", "40"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in BASE_KEYS + ["e2e", "cpu_baseline"]:
        assert k in d, k
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
