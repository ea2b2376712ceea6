"""CPU checks of the boundary: the library builds/loads, exports every declared symbol, host-side
logic works, and the product has no CPU path (compute entry points fail loudly without a GPU)."""
import os
import re

import numpy as np
import pytest

from tests.util import ROOT


def test_library_exports_every_declared_symbol():
    from laboetie_b200 import api, build
    build.build()
    L = api.load_library()
    hdr = open(os.path.join(ROOT, "include", "laboetie_gpu.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|const char\*)\s+(lbg_[a-z0-9_]+)\s*\(", hdr, re.M)))
    assert declared, "no declarations parsed"
    for s in declared:
        assert hasattr(L, s), f"{s} declared in include/laboetie_gpu.h but not exported"
    assert sorted(api.SYMBOLS) == declared
    assert L.lbg_abi_version() == 1
    assert L.lbg_status_string(1).decode().startswith("In equilibration, the population")
    assert L.lbg_status_string(2).decode() == "somewhere restpart is negative"


def test_partition_covers_lattice():
    from laboetie_b200 import api
    for lz, r in [(10, 3), (8, 8), (1024, 8), (7, 2), (5, 1)]:
        parts = [api.partition(lz, r, i) for i in range(r)]
        assert parts[0][0] == 0 and sum(p[1] for p in parts) == lz
        for a, b in zip(parts, parts[1:]):
            assert a[0] + a[1] == b[0]
        assert max(p[1] for p in parts) - min(p[1] for p in parts) <= 1
    from laboetie_b200.api import LbgError
    with pytest.raises(LbgError):
        api.partition(3, 4, 0)


def test_halo_plan_matches_velocity_table():
    from laboetie_b200 import api
    from oracle import oracle as O
    c = O.lbm_table()[0]
    up, down = api.halo_plan()
    assert sorted(up) == [l for l in range(19) if c[l][2] == 1]
    assert sorted(down) == [l for l in range(19) if c[l][2] == -1]
    # reference numbering (1-based): 6,12,13,16,17 up and 7,14,15,18,19 down
    assert [l + 1 for l in up] == [6, 12, 13, 16, 17] and [l + 1 for l in down] == [7, 14, 15, 18, 19]


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (it must never route through the oracle)."""
    import laboetie_b200 as lb
    from tests.conftest import _has_gpu
    if _has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(lb.LbgError) as e:
        lb.LaboetieGPU(np.zeros((4, 4, 4), np.int8))
    assert e.value.status == 10


def test_product_never_imports_oracle():
    """Nothing under laboetie_b200/ may import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "laboetie_b200")
    bad = re.compile(r"(^|\s)(import|from)\s+oracle\b|liblaboetie_oracle|oracle[/\\.](oracle|numpy_restatement|laboetie_oracle)|orc_[a-z_]+\(")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".f90", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not bad.search(txt), os.path.join(dp, f)


def test_compensating_force_field_matches_oracle():
    """Host-side mirror of equilibration.f90:388-487 against the oracle restatement."""
    from laboetie_b200 import driver
    from oracle import oracle as O
    for label, shape, pd in ((-1, (7, 5, 9), 3), (1, (5, 7, 9), 3), (-1, (5, 5, 5), 1), (1, (9, 9, 11), 5)):
        nat = O.geometry(label, *shape)
        f = [1e-3, -2e-3, 5e-4]
        a = O.compensate_force(nat, f, pd=pd, geometry_label=label)
        b = driver.compensating_force_field(nat, f, pd, None, label)
        assert a[3] == b[3]
        for x, y in zip(a[:3], b[:3]):
            assert np.array_equal(x, y)
        if label == -1:
            assert abs(a[0].sum()) < 1e-15
    with pytest.raises(ValueError):
        driver.compensating_force_field(O.geometry(-1, 4, 5, 5), [1, 0, 0], 1)
    with pytest.raises(ValueError):
        driver.compensating_force_field(O.geometry(-1, 5, 5, 5), [1, 0, 0], 2)
