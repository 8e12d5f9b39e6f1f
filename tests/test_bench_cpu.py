"""Host-side logic of bench.py that needs no GPU: the reference arm's JSON line, the clocks summary, the peaks file, the parity helper."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--batch", "4"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("resnet_18") and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_clocks_summary_and_peaks():
    c = bench.Clocks(0)
    rows = [["0", "1965", "1965", "640.5", "0x0", "Not Active", "Not Active", "Not Active", "Not Active"],
            ["0", "1950", "1965", "700.0", "0x4", "Not Active", "Not Active", "Not Active", "Active"],
            ["0", "1965", "1965", "650.0", "0x0", "Not Active", "Not Active", "Not Active", "Not Active"]]
    c.rows = [(float(i), r) for i, r in enumerate(rows)]
    s = c.summary(-1.0, 10.0)
    assert s["sm_mhz"] == 1965.0 and s["sm_max_mhz"] == 1965.0 and s["reasons"] == ["sw_power_cap"] and s["power_w_max"] == 700.0
    assert bench.Clocks(0).summary(0, 1)["reasons"] == ["nvidia-smi unavailable"]
    p = bench.peaks()
    assert p["hbm"] > 1000 and p["tf_sust"] <= p["tf_burst"] and p["src"] in ("measured", "fallback")


def test_mean3d_diff_helper():
    c = torch.load(os.path.join(ROOT, "tests", "golden", "backbone_cases.pt"))[0]
    ref = c["eval_uvd"][0]
    assert bench.mean3d_diff_mm(ref.clone(), ref, c["B"], c["J"], c["H"]) == 0.0
    assert 0.0 < bench.mean3d_diff_mm(ref + 1e-3, ref, c["B"], c["J"], c["H"]) < 0.05       # a 1e-3 UVD shift stays inside the 0.05 mm bound
