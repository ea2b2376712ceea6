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
    out = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", "cfg3", "--cpu-seconds", "3")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "MLUPS" and line["higher_is_better"] is True
    assert line["metric"].startswith("MLUPS") and line["value"] > 0 and line["dtype"] == "f64"
    assert line["steps"] == 2 and line["warmup"] == 1          # what was asked for is what was timed
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "slab of cfg3" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_4_threads"]["cores"] == 4 and line["cpu_4_threads"]["value"] > 0
    assert line["vs_baseline"] is None and line["config"]["workload"] == "cfg3"
    assert line["sample_is_whole_single_gpu_lattice"] is False and line["sample_lattice"][:2] == [256, 256]
    # ms_per_step is the time of one LB + one MP step of the sample
    n = line["sample_lattice"][0] * line["sample_lattice"][1] * line["sample_lattice"][2]
    assert abs(line["ms_per_step"] - n / line["value"] / 1e3) < 1e-6 * line["ms_per_step"]


def test_both_arms_print_the_same_config():
    """`--impl reference` runs "your arm's config": the dict is built by one function for both arms."""
    import argparse
    import bench
    args = argparse.Namespace(workload="cfg5w", check_every=1, in_place=False)
    c = bench.workload_config(args, 4)
    assert c["workload"] == "cfg5w" and c["lattice"] == [1024, 1024, 512] and c["parallelism"] == "z-slabs x4"
    assert c["phase_a_layout"] == "two-lattice" and "l2" in c and c["check_every"] == 1
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": workload_config(args, ') == 2


def test_reference_arm_whole_lattice():
    """--cpu-lattice full: the GPU arm's own single-GPU lattice, still K timed steps after W warm-up steps."""
    out = _run("--impl", "reference", "--steps", "3", "--warmup", "1", "--workload", "cfg2", "--cpu-lattice", "full")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["config"]["lattice"] == [64, 64, 256] and line["sample_lattice"] == [64, 64, 256]
    assert line["sample_is_whole_single_gpu_lattice"] is True and line["steps"] == 3 and line["warmup"] == 1
    assert "whole single-GPU lattice" in line["cpu_baseline"]["sample"]


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
