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


def test_host_side_index_helpers(tmp_path):
    """div_magic (exact multiply-shift division used by every kernel to turn a dense node index into
    (x, y, z)) and nbt_row (row numbering of the Phase-B neighbour table) are host-compilable: check
    them with g++ against plain integer division / the D3Q19 table."""
    import shutil
    import subprocess
    cuda_inc = "/usr/local/cuda/include"
    if not (shutil.which("g++") and os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h"))):
        pytest.skip("needs g++ and the CUDA headers")
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cstdint>
#include "lbg_internal.h"
int main() {
  long bad = 0;
  const int ds[] = {1, 2, 3, 5, 7, 31, 32, 33, 50, 64, 100, 255, 256, 257, 1000, 1023, 1024, 1025, 2500, 4096, 65535,
                    65536, 65537, 262144, 1048576, 1048577, 3000000, 16777216, 100000007, 1073741824, 2147483647};
  for (int d : ds) {
    uint32_t m; int sh;
    lbg::div_magic(d, &m, &sh);
    auto chk = [&](uint32_t g) { if ((uint32_t)(((unsigned long long)g * m) >> sh) != g / (uint32_t)d) ++bad; };
    for (uint32_t g = 0; g < 300000; ++g) chk(g);
    for (uint32_t g = 2147483647u; g > 2147483647u - 300000; --g) chk(g);
    for (unsigned long long k = 1; k < 300000; ++k) {
      const unsigned long long g = k * (unsigned)d;
      if (g >= (1ull << 31)) break;
      chk((uint32_t)g); chk((uint32_t)g - 1);
    }
    for (uint32_t g = 1; g < (1u << 31); g += 7919) chk(g);
  }
  // every direction with cx == 0 except the rest one is the centre of exactly one of the 8 rows
  int seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int l = 1; l < d3q19::NV; ++l) {
    const int r = lbg::nbt_row(d3q19::cy(l), d3q19::cz(l));
    if (d3q19::cy(l) == 0 && d3q19::cz(l) == 0) { if (r != -1) ++bad; continue; }
    if (r < 0 || r > 7) { ++bad; continue; }
    if (d3q19::cx(l) == 0) ++seen[r];
    // a row is shared by a direction and by the one with the same (cy, cz) only
    for (int m = 1; m < d3q19::NV; ++m)
      if ((lbg::nbt_row(d3q19::cy(m), d3q19::cz(m)) == r) != (d3q19::cy(m) == d3q19::cy(l) && d3q19::cz(m) == d3q19::cz(l))) ++bad;
  }
  for (int r = 0; r < 8; ++r) if (seen[r] != 1) ++bad;
  // 16-bit row-centre deltas of the packed table: exact round trip over the whole range, both flag values
  for (int d = -lbg::NBT_DELTA_BIAS; d < lbg::NBT_DELTA_BIAS; ++d)
    for (int f = 0; f < 2; ++f) {
      const uint32_t h = lbg::nbt_enc16(d, f != 0);
      if (h > 0xffffu || lbg::nbt_dec16(h) != d || (int)((h >> 15) & 1u) != f || !lbg::nbt_delta_fits(d)) ++bad;
      if (lbg::nbt_dec16((h << 16) >> 16) != d) ++bad;
    }
  if (lbg::nbt_delta_fits(lbg::NBT_DELTA_BIAS) || lbg::nbt_delta_fits(-lbg::NBT_DELTA_BIAS - 1)) ++bad;
  printf("%ld\n", bad);
  return bad != 0;
}
''')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-D__host__=", "-D__device__=", "-I", cuda_inc,
                           "-I", os.path.join(ROOT, "laboetie_b200", "csrc"), "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "0", out.stdout
