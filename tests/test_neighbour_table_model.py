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
    # packing (lbg_internal.h): rows (0,+-z) absolute, the others 16-bit deltas (+16384, bit 15 = centre fluid)
    # from the own fid (rows 0, 1), from row 2 (rows 4, 5) and from row 3 (rows 6, 7); out of range -> slow
    BIAS = 16384
    ref_of = {0: fid, 1: fid, 4: centre_c[2], 5: centre_c[2], 6: centre_c[3], 7: centre_c[3]}
    enc = {}
    for r, ref in ref_of.items():
        d = centre_c[r].astype(np.int64) - ref.astype(np.int64)
        slow |= (d < -BIAS) | (d >= BIAS)
        enc[r] = (((d + BIAS) & 0x7fff) | (centre_f[r].astype(np.int64) << 15)).astype(np.uint32)
    w = [centre_c[2].astype(np.uint32) | (centre_f[2].astype(np.uint32) << 31),
         centre_c[3].astype(np.uint32) | (centre_f[3].astype(np.uint32) << 31),
         enc[0] | (enc[1] << 16), enc[4] | (enc[5] << 16), enc[6] | (enc[7] << 16)]
    # decode, as mp_step_kernel does
    dec16 = lambda h: (h & 0x7fff).astype(np.int64) - BIAS     # noqa: E731
    c2, c3 = (w[0] & 0x3fffffff).astype(np.int64), (w[1] & 0x3fffffff).astype(np.int64)
    dc = {2: c2, 3: c3, 0: fid + dec16(w[2]), 1: fid + dec16(w[2] >> 16), 4: c2 + dec16(w[3]), 5: c2 + dec16(w[3] >> 16),
          6: c3 + dec16(w[4]), 7: c3 + dec16(w[4] >> 16)}
    df = {2: w[0] >> 31, 3: w[1] >> 31, 0: (w[2] >> 15) & 1, 1: w[2] >> 31, 4: (w[3] >> 15) & 1, 5: w[3] >> 31,
          6: (w[4] >> 15) & 1, 7: w[4] >> 31}
    for r in range(8):
        ok = fluid & ~slow
        assert np.array_equal(dc[r][ok], centre_c[r][ok]) and np.array_equal(df[r][ok].astype(bool), centre_f[r][ok])
    centre_c = {r: dc[r] for r in range(8)}
    centre_f = {r: df[r].astype(np.int64) for r in range(8)}
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


def regular_nodes(nat):
    """numpy statement of the "regular" bit mp_init_kernel records on the rank-lookup path: every fluid neighbour's
    fluid id equals the node's own id plus the neighbour's dense offset; a solid neighbour's id + offset is in range."""
    lz, ly, lx = nat.shape
    fluid = (nat == 0)
    flat = fluid.ravel()
    rank = np.concatenate([[0], np.cumsum(flat)])[:-1]
    nf = int(flat.sum())
    nfa = max(32, (nf + 31) // 32 * 32)
    z, y, x = np.meshgrid(np.arange(lz), np.arange(ly), np.arange(lx), indexing="ij")
    dense = lambda zz, yy, xx: ((zz % lz) * ly + (yy % ly)) * lx + (xx % lx)   # noqa: E731
    own = dense(z, y, x)
    fid = rank[own].astype(np.int64)
    reg = fluid.copy()
    for l in range(1, 19):
        cx, cy, cz = C[l]
        g = dense(z + cz, y + cy, x + cx)
        guess = fid + (g - own)                       # own id + dense offset (periodic wraps included)
        nb_fluid = flat[g]
        reg &= np.where(nb_fluid, rank[g] == guess, (guess >= 0) & (guess < nfa))
    return reg, fluid


def test_regular_nodes_of_open_and_porous_geometries():
    """A slit is regular everywhere except in the planes whose solid neighbours would index outside the arrays; a
    random porous lattice has hardly any regular node.  (Exactness of the rule itself is what the GPU parity tests of the
    rank-lookup path check: a regular node's arithmetic ids must reproduce the oracle's P bit for bit.)"""
    nat = O.geometry(1, 16, 12, 20)                  # slit: planes z = 0 and z = lz-1 solid
    reg, fluid = regular_nodes(nat)
    assert not reg[~fluid].any()
    assert reg[2:-2].all()                           # interior planes: all regular
    assert reg[fluid].mean() > 0.85
    bulk = O.geometry(-1, 9, 7, 5)                   # all fluid, periodic: every node regular
    assert regular_nodes(bulk)[0].all()
    reg, fluid = regular_nodes(random_nature(24, 12, 9, 0.3, 11))
    assert reg[fluid].mean() < 0.05
