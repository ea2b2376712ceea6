"""Host model of the Phase-B neighbour table (laboetie_b200/csrc/lbg_internal.h NBT_*, mp_kernels.cu).

No GPU: a numpy restatement of what mp_init_kernel stores (per fluid node and neighbouring row: the rank
position of the row's centre node and a "centre is fluid" bit; a "slow" flag on the periodic x seam or
where a derived index would leave the arrays) and of how mp_step_kernel decodes it (fid(x+1) = c +
centre_fluid, fid(x-1) = c - 1).  Checked against direct periodic neighbour look-ups on random lattices:
every FLUID neighbour of every non-slow node must decode to exactly its fluid id, every decoded index must
be in range, and every node on the x seam must be flagged slow.  The GPU parity tests check the kernels
themselves (tests/test_gpu_parity.py, both paths)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import random_nature

C = np.array(O.lbm_table()[0])          # (19, 3): cx, cy, cz, reference order


def nbt_row(cy, cz):
    if cz == 0:
        return 0 if cy > 0 else (1 if cy < 0 else -1)
    if cz > 0:
        return 2 if cy == 0 else (4 if cy > 0 else 5)
    return 3 if cy == 0 else (6 if cy > 0 else 7)


def build_and_check(nat):
    lz, ly, lx = nat.shape
    fluid = (nat == 0)
    flat = fluid.ravel()                                  # dense order: x fastest, then y, then z
    rank = np.concatenate([[0], np.cumsum(flat)])[:-1]    # number of fluid nodes before each dense node
    nf = int(flat.sum())
    nfa = max(32, (nf + 31) // 32 * 32)
    z, y, x = np.meshgrid(np.arange(lz), np.arange(ly), np.arange(lx), indexing="ij")
    dense = lambda zz, yy, xx: ((zz % lz) * ly + (yy % ly)) * lx + (xx % lx)   # noqa: E731
    own = dense(z, y, x)
    fid = rank[own]
    n_checked = 0
    # table: centre rank position / centre-fluid bit per row, slow flag
    centre_c, centre_f = {}, {}
    slow = (x == 0) | (x == lx - 1) | (fid < 1) | (fid + 1 >= nfa)
    for l in range(1, 19):
        cx, cy, cz = C[l]
        if cx == 0:
            r = nbt_row(cy, cz)
            g = dense(z + cz, y + cy, x)
            centre_c[r], centre_f[r] = rank[g], flat[g]
            slow |= (rank[g] < 1) | (rank[g] + 1 >= nfa)
    assert sorted(centre_c) == list(range(8))
    for l in range(1, 19):
        cx, cy, cz = C[l]
        r = nbt_row(cy, cz)
        if r < 0:
            dec = fid + cx
        elif cx == 0:
            dec = centre_c[r]
        elif cx > 0:
            dec = centre_c[r] + centre_f[r]
        else:
            dec = centre_c[r] - 1
        g = dense(z + cz, y + cy, x + cx)
        ok_nodes = fluid & ~slow
        assert (dec[ok_nodes] >= 0).all() and (dec[ok_nodes] < nfa).all()
        sel = ok_nodes & flat[g].reshape(nat.shape)       # fluid node, not slow, fluid neighbour
        assert np.array_equal(dec[sel], rank[g][sel]), f"direction {l}"
        n_checked += int(sel.sum())
    # every fluid node on the periodic x seam takes the rank-lookup path
    assert slow[fluid & ((x == 0) | (x == lx - 1))].all()
    return n_checked


@pytest.mark.parametrize("shape,p,seed", [((7, 6, 40), 0.3, 1), ((5, 9, 33), 0.6, 2), ((4, 4, 4), 0.5, 3), ((3, 2, 70), 0.2, 4),
                                          ((9, 1, 12), 0.25, 5), ((1, 12, 9), 0.2, 6), ((6, 5, 1), 0.3, 7), ((2, 3, 2), 0.3, 8)])
def test_decoded_neighbours_equal_lookups(shape, p, seed):
    lz, ly, lx = shape
    nat = random_nature(lx, ly, lz, p, seed)
    n = build_and_check(nat)
    if lx >= 8:
        assert n > 0


def test_reference_geometries():
    for nat in (O.geometry(1, 8, 8, 16), O.geometry(2, 11, 11, 4), O.geometry(3, 12, 12, 12)):
        assert build_and_check(nat) > 0
