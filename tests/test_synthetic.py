"""The benchmark geometry builders (laboetie_b200/synthetic.py) against the literal oracle builders."""
import numpy as np
import pytest

from laboetie_b200 import synthetic as S
from oracle import oracle as O


@pytest.mark.parametrize("lx", [3, 5, 8, 11, 26, 51, 64])
def test_cylinder_matches_reference_builder(lx):
    assert np.array_equal(S.cylinder(lx, lx, 3), O.geometry(2, lx, lx, 3))


@pytest.mark.parametrize("lx", [4, 6, 9, 16, 33])
def test_bcc_matches_reference_builder(lx):
    assert np.array_equal(S.bcc(lx, lx, lx), O.geometry(3, lx, lx, lx))


def test_slit_matches_reference_builder():
    assert np.array_equal(S.slit(4, 3, 9), O.geometry(1, 4, 3, 9))


@pytest.mark.parametrize("builder,kw", [(S.porous_spheres, dict(radius=3)), (S.bernoulli, {}), (S.bcc, {}), (S.slit, {})])
def test_slabs_are_windows_of_the_global_geometry(builder, kw):
    lx = ly = lz = 24
    full = builder(lx, ly, lz, **kw)
    for k0, nzl in [(0, 24), (0, 7), (7, 9), (16, 8)]:
        slab = builder(lx, ly, lz, k0=k0 - 1, nz=nzl + 2, **kw)
        idx = np.arange(k0 - 1, k0 + nzl + 1) % lz
        assert np.array_equal(slab, full[idx])


def test_porous_spheres_porosity_and_determinism():
    a = S.porous_spheres(64, 64, 48, radius=4)
    assert abs((1 - a.mean()) - 0.6) < 0.02
    assert np.array_equal(a, S.porous_spheres(64, 64, 48, radius=4))
    b = S.bernoulli(32, 32, 32)
    assert abs(b.mean() - 0.25) < 0.01
