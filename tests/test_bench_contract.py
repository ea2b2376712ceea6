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
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "cfg3", "--cpu-lattice", "crop")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "MLUPS" and line["higher_is_better"] is True
    assert line["metric"].startswith("MLUPS") and line["value"] > 0 and line["dtype"] == "f64"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "crop" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_4_threads"]["cores"] == 4 and line["cpu_4_threads"]["value"] > 0
    assert line["vs_baseline"] is None and line["config"]["workload"] == "cfg3"
    assert line["config"]["whole_lattice_of_the_gpu_arm_at_n1"] is False


def test_reference_arm_whole_lattice_and_step_count():
    """With enough host memory the reference arm runs the GPU arm's own lattice (same_config), and `steps` is the
    number of steps it really timed, whatever --steps asked for."""
    out = _run("--impl", "reference", "--steps", "20", "--warmup", "5", "--workload", "cfg2", "--cpu-lattice", "full")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["config"]["lattice"] == [64, 64, 256] and line["config"]["whole_lattice_of_the_gpu_arm_at_n1"] is True
    assert line["steps"] == 1 and line["requested_steps"] == 20 and line["warmup"] == 0
    assert "whole lattice" in line["cpu_baseline"]["sample"]
    assert abs(line["ms_per_step"] - 64 * 64 * 256 / line["value"] / 1e3) < 1e-6 * line["ms_per_step"]


def test_traffic_file_is_tied_to_the_kernel_sources():
    import bench
    h = bench.kernel_source_hash()
    assert len(h) == 16 and h == bench.kernel_source_hash()
    tp = os.path.join(ROOT, "profiles", "traffic_cfg5w.json")
    if os.path.exists(tp):
        tr = json.load(open(tp))
        assert "kernel_source_hash" in tr, "profiles/traffic_cfg5w.json must say which kernel sources it was measured on"


def test_product_arm_has_no_cpu_path():
    from tests.conftest import _has_gpu
    if _has_gpu():
        pytest.skip("GPU present")
    out = _run("--steps", "1", "--warmup", "3", "--workload", "cfg2", "--no-cpu-baseline", "--no-e2e", "--also", "")
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)
