"""bench.py's JSON-line contract (CPU side): the reference arm runs here, and the committed round artefact of the
GPU arm carries every key the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--lines", "2000", "--shells", "5",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["unit"] == "packets/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_committed_gpu_bench_line_has_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "r02_bench_final_default_1e8.json")
    if not os.path.exists(path):
        pytest.skip("the round's final bench line is not committed yet")
    d = json.loads(open(path).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"} <= set(d)
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["vs_baseline"] is None
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 0 and d["warmup"] >= 3
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # a timed step pays for its fresh packets: seed expansion + three ordering kernels + transport + two epilogues
    assert d["gpu_launches"] >= 7 * d["steps"] and e["steps"] == d["steps"]
    # round 2: the BASELINE legs ride in the same line, each with its own e2e / roofline / cpu_baseline / parity
    for k in ("2", "4", "5"):
        leg = d["configs"][k]
        assert leg["value"] > 0 and leg["e2e"]["value"] > 0 and leg["warmup"] >= 3 and leg["scaling"] == "strong"
        assert leg["cpu_baseline"]["value"] > 0 and leg["roofline"]["kernel_ms"] > 0
        p = leg["parity"]
        assert p["counters_equal"] and p["spectrum_l2_vs_oracle"] < 1e-10 and max(p["max_rel_err"].values()) < 1e-10
    assert d["configs"]["4"]["parity"]["virtual_spectrum_l2_vs_oracle"] < 1e-10
    assert d["configs"]["5"]["parity"]["photo_ion_statistics_equal"]
    assert d["parity"]["counters_equal"] and d["parity"]["fused_spectrum_l2_vs_oracle"] < 1e-10
    assert r["traffic"] is not None and r["issue_active_pct"] is not None
    # per-iteration table preparation either side of the path (§8f ranks 3 / 4) rides in the same line
    t = d["tables"]
    assert "error" not in t and t["host_tables"]["ms"] > 0 and t["device_tables"]["ms"] > 0 and t["source_function"]["sweeps"] > 0
