"""CPU: `bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints ONE JSON line carrying the
contract's keys, on a small sample so that it runs in seconds.  The arm times the oracle port -- it is the one place
besides tests/ and smoke() that may execute oracle/ -- and launches nothing on a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2",
                          "--warmup", "1", "--cpu-sample-log-n", "12"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    assert line["metric"] == "BLS12-381 MSM scalar-mults/sec at 2^26" and line["unit"] == "scalar-mults/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] > 0 and abs(line["value"] - line["cpu_baseline"]["value"]) < 1e-6 * line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None
    assert "2^12" in line["config"]["sample"]
