"""The committed evidence must agree with itself (CPU): the ncu launch list of the bench command holds one
full-batch launch of the headline kernel per step, their durations agree with the CUDA-event time `bench.py`
reported in the same code state, and the roofline fields of the bench line are consistent with each other."""
import csv
import json
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PROF = ROOT / "profiles"


def _launches():
    rows = list(csv.reader(open(PROF / "r1_bench_launches.csv")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    col = {h: i for i, h in enumerate(rows[hdr])}
    out = []
    for r in rows[hdr + 1:]:
        if len(r) > col["Metric Value"] and r[col["Metric Name"]] == "gpu__time_duration.sum":
            out.append((r[col["Kernel Name"]], r[col["Grid Size"]], float(r[col["Metric Value"]].replace(",", ""))))
    return out


def test_launch_list_matches_bench_line():
    line = json.loads((PROF / "r1_bench_n1.json").read_text().strip().splitlines()[-1])
    roof = line["roofline"]
    head = [ns for name, grid, ns in _launches() if re.search(r"patch_gather_kernel<2, 0, 1, 0, 0, 16>", name)]
    full = [ns for ns in head if ns > 0.5 * max(head)]
    # bench command of the list: --steps 5 --warmup 3 -> 8 full-batch launches, one per step
    assert len(full) == 8
    assert max(full) / min(full) < 1.02
    ncu_ms = sum(full) / len(full) / 1e6
    assert abs(ncu_ms / roof["kernel_ms"] - 1.0) < 0.05          # cold-cache, serialised vs CUDA events: same kernel time
    assert line["gpu_launches"] == line["steps"]                  # one launch per timed step
    assert abs(line["ms_per_step"] - roof["kernel_ms"]) < 1e-9    # the step IS that launch


def test_roofline_fields_are_consistent():
    line = json.loads((PROF / "r1_bench_n1.json").read_text().strip().splitlines()[-1])
    roof = line["roofline"]
    achieved = roof["algorithmic_bytes_per_launch"] / (roof["kernel_ms"] * 1e-3) / 1e9
    assert abs(achieved / roof["achieved"] - 1.0) < 1e-6
    assert abs(roof["achieved"] / roof["peak"] - roof["frac"]) < 1e-9
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s"
    # measured DRAM traffic of the launch: within 5 % above the algorithmic bytes (no wasted re-reads)
    assert 1.0 <= roof["traffic"] / roof["algorithmic_bytes_per_launch"] < 1.05
    # 16 nelem + 32 nnode bytes per problem (+ the mesh once)
    nelem, nnode, B = 999941, 578 * 578, 512
    assert roof["algorithmic_bytes_per_launch"] == B * (16 * nelem + 32 * nnode) + 8 * nelem + 16 * nnode
    assert abs(line["value"] - B * nelem / (line["ms_per_step"] * 1e-3)) / line["value"] < 1e-9
    for key in ("cpu_baseline", "e2e", "clocks"):
        assert key in line
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
