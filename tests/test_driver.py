"""The compiled driver mirror (laboetie_b200/driver): input parsing on CPU, a full run on the GPU."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests.util import GOLDEN, ROOT

DRV = os.path.join(ROOT, "laboetie_b200", "driver")
EXE = os.path.join(DRV, "laboetie_driver")


def _build():
    from laboetie_b200 import build
    build.build()
    subprocess.check_call(["make", "-C", DRV, "-s"])


def test_driver_parses_stock_input():
    _build()
    out = subprocess.run([EXE, "--input", os.path.join(GOLDEN, "lb.in.stock"), "--check-input"], stdout=subprocess.PIPE,
                         text=True, check=True).stdout
    assert "1 x 1 x 102" in out and "geometryLabel 1" in out and "2 solid nodes" in out


def test_driver_reads_geom_in_and_pbm(tmp_path):
    _build()
    shutil.copy(os.path.join(GOLDEN, "geom.in_chromat_1disks-dia10-1x50x50_v1"), tmp_path / "geom.in")
    (tmp_path / "lb.in").write_text("lx = 1\nly = 50\nlz = 50\ngeometryLabel = 0 # custom\nf_ext = 0.0 1.e-5 0.0\n")
    out = subprocess.run([EXE, "--input", str(tmp_path / "lb.in"), "--check-input"], stdout=subprocess.PIPE, text=True,
                         check=True).stdout
    assert "79 solid nodes" in out and "f_ext = 0 1e-05 0" in out
    shutil.copy(os.path.join(GOLDEN, "geom.pbm"), tmp_path / "geom.pbm")
    (tmp_path / "lb.in").write_text("lx = 1\nly = 91\nlz = 25\ngeometryLabel = 11\n")
    out = subprocess.run([EXE, "--input", str(tmp_path / "lb.in"), "--check-input"], stdout=subprocess.PIPE, text=True,
                         check=True).stdout
    from oracle import oracle as O
    n = int(O.read_pbm(os.path.join(GOLDEN, "geom.pbm"), 1, 91, 25).sum())
    assert f"{n} solid nodes" in out


def test_driver_rejects_what_the_reference_rejects(tmp_path):
    _build()
    (tmp_path / "lb.in").write_text("lx = 4\nly = 4\nlz = 4\ngeometryLabel = 1\nsigma = 0.1\n")
    r = subprocess.run([EXE, "--input", str(tmp_path / "lb.in"), "--check-input"], stderr=subprocess.PIPE, text=True)
    assert r.returncode != 0 and "uncharged" in r.stderr


@pytest.mark.gpu
def test_driver_runs_config1_end_to_end(tmp_path):
    """BASELINE config 1 through the compiled driver: files as the reference writes them, values as the oracle."""
    _build()
    shutil.copy(os.path.join(GOLDEN, "geom.in_chromat_1disks-dia10-1x50x50_v1"), tmp_path / "geom.in")
    (tmp_path / "lb.in").write_text(
        "lx = 1\nly = 50\nlz = 50\ngeometryLabel = 0\nf_ext = 0.0 1.e-5 0.0\n"
        "tracer_Db = 0.01\ntracer_ka = 0.1\ntracer_kd = 0.01\nmaximum_moment_propagation_steps = 400\n")
    subprocess.run([EXE, "--input", str(tmp_path / "lb.in"), "--outdir", str(tmp_path / "output"), "--quiet"], check=True)
    g = np.load(os.path.join(GOLDEN, "tuto_cfg1_oracle.npz"))
    l2 = np.loadtxt(tmp_path / "output" / "l2err.dat")
    assert l2.shape[0] == int(g["t_exit"]) and np.array_equal(l2[:, 1], g["l2err"])
    vacf = np.loadtxt(tmp_path / "output" / "vacf.dat")
    assert vacf.shape[0] == 401
    assert np.allclose(vacf[:, 1:], g["vacf"], rtol=1e-12, atol=1e-12 * np.abs(g["vacf"]).max())
    f2 = np.loadtxt(tmp_path / "output" / "mass-flux_field_2d_at_x.eq.1.dat")
    assert np.array_equal(f2[:, 2].reshape(50, 50), g["jy"][:, :, 0].T)
    prof = [l for l in open(tmp_path / "output" / "mass-flux_profile_along_z.dat") if not l.startswith("#") and l.strip()]
    last = np.array([[float(x) for x in l.split()] for l in prof[-50:]])
    assert np.allclose(last[:, 1:], g["prof_z"][:, :3], rtol=1e-12, atol=1e-300)


@pytest.mark.gpu
def test_driver_compensate_f_ext(tmp_path):
    """compensate_f_ext = T through the compiled driver against the Python mirror (same C ABI underneath)."""
    _build()
    import laboetie_b200 as lb
    from laboetie_b200 import driver
    from oracle import oracle as O
    (tmp_path / "lb.in").write_text(
        "lx = 7\nly = 5\nlz = 9\ngeometryLabel = -1\nf_ext = 1.e-4 0.0 3.e-4\ncompensate_f_ext = T\n"
        "dominika_particle_diameter = 3\nrelaxation_time = 0.9\ntarget_error = 1.e-9\n")
    subprocess.run([EXE, "--input", str(tmp_path / "lb.in"), "--outdir", str(tmp_path / "output"), "--quiet"], check=True)
    nat = O.geometry(-1, 7, 5, 9)
    with lb.LaboetieGPU(nat) as sim:
        r = driver.equilibration_compensated(sim, nat, [1e-4, 0, 3e-4], tau=0.9, target_error=1e-9,
                                             particle_diameter=3, geometry_label=-1)
    l2 = np.loadtxt(tmp_path / "output" / "l2err.dat")
    assert l2.shape[0] == r["t_exit"] and np.array_equal(l2[:, 1], r["l2err"])
    v = np.loadtxt(tmp_path / "output" / "v_centralnode.dat")
    assert np.array_equal(v, r["v_centralnode"])
