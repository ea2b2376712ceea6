"""One process per GPU, launched by torchrun: the path bench.py --gpus N and every multi-GPU user runs
(cudaIpcOpenMemHandle-mapped neighbour lattices and mailboxes; tests/test_multigpu.py drives its ranks as threads of
one process and therefore takes the plain peer-access branch).  Needs >= 2 GPUs on the box.
The single-GPU variant of the same check (bench.verify_launch) runs on any GPU box."""
import json
import os
import subprocess
import sys

import pytest

from tests.util import ROOT

pytestmark = pytest.mark.gpu


def _ndev():
    import ctypes
    from laboetie_b200 import api
    n = ctypes.c_int()
    api.load_library().lbg_device_count(ctypes.byref(n))
    return n.value


@pytest.mark.parametrize("in_place", [False, True])
def test_verify_launch_single_gpu(in_place):
    import bench
    v = bench.verify_launch(None, 0, 1, 0, in_place=in_place)
    assert v["ok"] and v["max_abs_diff"] == 0.0 and v["path"] == "single GPU", v


@pytest.mark.parametrize("nranks", [2, 4])
def test_process_per_gpu_matches_oracle(nranks):
    if _ndev() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    port = 29500 + os.getpid() % 500
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp_verify_worker.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["ok"] and line["max_abs_diff"] == 0.0 and line["path"] == "ipc" and line["ranks"] == nranks, line
    assert line["second_run_ok"] and line["in_place_ok"], line
