"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line (the CPU
restatement of the reference on a bounded crop), and the product arm refuses to run without a CUDA device
instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

from tests.util import ROOT


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_json_line():
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "cfg3")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "MLUPS" and line["higher_is_better"] is True
    assert line["metric"].startswith("MLUPS") and line["value"] > 0 and line["dtype"] == "f64"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "crop" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_4_threads"]["cores"] == 4 and line["cpu_4_threads"]["value"] > 0
    assert line["vs_baseline"] is None and line["config"]["workload"] == "cfg3"


def test_product_arm_has_no_cpu_path():
    from tests.conftest import _has_gpu
    if _has_gpu():
        pytest.skip("GPU present")
    out = _run("--steps", "1", "--warmup", "3", "--workload", "cfg2", "--no-cpu-baseline", "--no-e2e", "--also", "")
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)
