"""The committed measurement files under profiles/ are evidence: check that the bench lines obey bench.py's contract and that the
numbers quoted from them are arithmetically consistent (a line edited by hand, or produced by a bench that dropped a key, fails
here).  No GPU, no oracle: file checks only."""
import csv
import glob
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "roofline", "gpu_launches", "clocks", "e2e"]


def lines():
    out = []
    for f in sorted(glob.glob(os.path.join(P, "r2_bench_n1*.json")) + glob.glob(os.path.join(P, "r2_scale_N*_c*.json"))):
        with open(f) as fh:
            out.append((os.path.basename(f), json.load(fh)))
    return out


@pytest.mark.parametrize("name,d", lines(), ids=[n for n, _ in lines()])
def test_bench_line_contract(name, d):
    if d.get("impl") == "reference":
        assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["value"] == d["value"] and d["gpu_launches"] == 0
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
        return
    for k in REQUIRED:
        assert k in d, k
    assert d["metric"] == "geometries/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    m = re.search(r"_N(\d+)_", name)
    assert d["n_gpus"] == (int(m.group(1)) if m else 1)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (d["config"]["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert d["config"]["kernel_ms"] <= d["ms_per_step"] * 1.0001  # the kernel is part of the step
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]  # copies inside: slower than resident
    v = d.get("verify")
    assert v, "a bench line without its post-run verification"
    assert all(val is True for key, val in v.items() if key.endswith("equal_oracle") or key.endswith("bit_identical") or key.endswith("bit_exact")
               or key.endswith("equal_bincount") or key.startswith("distance_within"))
    assert not set(d["clocks"].get("reasons", [])) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_join_arms_process_the_points_they_claim():
    for name, d in lines():
        if d.get("impl") == "reference":
            continue
        w = d["config"]["workload"]
        m = re.match(r"(\d+) random points per GPU", w)
        if not m:
            continue
        per_gpu = int(m.group(1))
        assert abs(d["value"] * d["ms_per_step"] * 1e-3 - per_gpu * d["n_gpus"]) < 1e-6 * per_gpu * d["n_gpus"], name
        assert d["roofline"]["algorithmic_bytes_per_launch"] >= 16 * per_gpu  # 16 B per point (+ the polygon side once)


def test_launch_list_share_agrees_with_the_live_step():
    """B200_PROFILING.md: ncu's per-launch times are cold and serialised, so the kernel's SHARE of the step must agree"""
    d = json.load(open(os.path.join(P, "r2_bench_n1.json")))
    live = d["config"]["kernel_ms"] / d["ms_per_step"]
    share = None
    for ln in open(os.path.join(P, "r2_launches_summary.txt")):
        if "k_pip_stream" in ln:
            share = float(re.search(r"share=([0-9.]+)", ln).group(1))
    assert share is not None and abs(share - live) < 0.08, (share, live)


def test_traffic_file_matches_the_ncu_summary():
    t = json.load(open(os.path.join(P, "r2_pip_traffic.json")))
    rows = {r[0]: r for r in csv.reader(open(os.path.join(P, "r2_pip_stream_summary.csv")))}
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(rows["dram__bytes_read.sum"][2]) * unit[rows["dram__bytes_read.sum"][1]]
    wr = float(rows["dram__bytes_write.sum"][2]) * unit[rows["dram__bytes_write.sum"][1]]
    assert abs(t["dram_bytes_per_launch"] - (rd + wr)) < 1e-6 * (rd + wr)
    assert "k_pip_stream" in t["kernel"] and "k_pip_stream" in rows["kernel"][2]
    # traffic close to the algorithmic bytes (read 1.61 GB + ids 0.4 GB): no wasted re-reads
    assert 0.95 < t["dram_bytes_per_launch"] / (1.6104e9 + 0.4e9) < 1.15
