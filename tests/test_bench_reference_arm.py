"""bench.py --impl reference (the CPU arm the driver runs beside the GPU arm) on a small mesh: the JSON line's contract."""
import json
import os
import subprocess
import sys

from common import ROOT


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--ref-size", "32", "--ref-procs", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "alpha-advection cell-updates/sec" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    # the reference's own class when oracle/_ref/libref_solver.so is there (the port's figure rides along), else the port
    have_class = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_solver.so"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_class else "port")
    assert d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    if have_class:
        assert d["cpu_baseline"]["port_value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_port_only():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-size", "24", "--ref-procs", "2", "--ref-kind", "port"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["kind"] == "port" and "port_value" not in d["cpu_baseline"]
