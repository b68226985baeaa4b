"""bench.py --impl reference (the driver's reference arm): one JSON line with the contract's keys, on the port alone and — when this image
has the Mesa libGL — with the reference's shaders on llvmpipe in front of it. No GPU involved."""
import json
import os
import subprocess
import sys

import pytest

import glref_util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_arm(*extra):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *extra],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["workload"].startswith("C3: fixed batch of 1024 scenes x 20 objects")
    return line


def test_reference_arm_port_only():
    line = run_arm("--ref-arm", "port")
    assert line["cpu_baseline"]["kind"] == "port"


@pytest.mark.skipif(glref_util.available() is not None, reason=str(glref_util.available()))
def test_reference_arm_runs_the_reference_shaders_on_llvmpipe():
    line = run_arm()
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and "oracle/_ref/glref" in cb["sample"]
    assert cb["port"]["value"] > 0 and cb["port"]["unit"] == "frames/s"         # the port timed beside it, on the same scenes
