"""CPU checks of the measurement contract's host side: the committed DRAM-traffic figures bench.py reports, the reference
arm's JSON line, and the product arm's refusal to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_traffic_figures_are_tied_to_the_workload_they_were_captured_on():
    import bench
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
        t = json.load(fh)
    for cfg in ("c1", "c2", "c3", "c4", "c5"):
        e = t[cfg]
        assert os.path.exists(os.path.join(ROOT, e["source"])), e["source"]
        assert e["bytes"] >= e["algorithmic_bytes"] > 0                       # a kernel cannot read less than its input
        assert bench.known_traffic(cfg, e["algorithmic_bytes"]) == e["bytes"]
        assert bench.known_traffic(cfg, e["algorithmic_bytes"] // 2) is None   # another size: no figure rather than a wrong one
    assert bench.known_traffic("nope", 1) is None


def test_reference_arm_prints_the_contract_line_and_product_arm_needs_a_device():
    import torch
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-seconds", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "input_GBps" and line["unit"] == "GB/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)
