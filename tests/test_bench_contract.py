"""CPU tier: the JSON line bench.py prints keeps the driver's contract.  Checked on the line committed from the last
B200 run of the round (profiles/r2_bench_line.json) and on the argument surface of bench.py itself."""
import json
import os
import subprocess
import sys

from conftest import ROOT

LINE = os.path.join(ROOT, "profiles", "r2_bench_line.json")


def test_committed_bench_line_has_every_contract_key():
    d = json.load(open(LINE))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["value"] > 0 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "l2" in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1
    assert {"frac_step", "achieved_step"} <= set(r) and r["traffic"] is None or r["traffic"] > 0
    assert d["config"]["ddc"] in ("poly", "exact") and "exact_mode" in d and "occupancy_sweep" in d
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_bench_argument_surface():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload", "--ddc", "--input"):
        assert flag in out.stdout
