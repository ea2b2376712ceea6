"""Host model of the in-place (AA) Lattice-Boltzmann step ACROSS z-slabs.

The CUDA library runs the AA pattern on one slab only so far.  This numpy model fixes the multi-slab
protocol before it is built: which arrays travel after the local "even" step (forward: into the
neighbours' halo planes), which travel back after the pull/push "odd" step (reverse: from my halo
planes into the neighbours' own boundary planes), and why the return trip must be MASKED -- a halo
slot whose owner node (one plane inside the sender) is solid was never written by the sender and
still holds the stale forward copy, while the receiver has written its bounce-back value there.
Results must equal the single-domain restatement bit for bit after every step; unexchanged halo
arrays are poisoned with NaN so that any read outside the exchanged lists shows up.
"""
import numpy as np
import pytest

from oracle import numpy_restatement as R
from tests.util import random_nature

C, INV = R.C, R.INV
UP = [l for l in range(19) if C[l][2] == 1]      # cz = +1
DOWN = [l for l in range(19) if C[l][2] == -1]   # cz = -1


class Slab:
    def __init__(self, nat, k0, nzl):
        lz = nat.shape[0]
        self.k0, self.nzl = k0, nzl
        planes = [(k0 - 1 + p) % lz for p in range(nzl + 2)]
        self.nat = nat[planes]                       # (nzl+2, ly, lx) with halo planes
        self.fluid = self.nat == 0
        self.F = np.full((19,) + self.nat.shape, np.nan)
        self.own = slice(1, nzl + 1)

    def shifted(self, a, c):
        """a at r + c for the own planes: x, y periodic, z plain (the halo planes stand in)."""
        b = np.roll(a, shift=(-c[1], -c[0]), axis=(1, 2))
        return b[1 + c[2]: 1 + c[2] + self.nzl]


def even_step(s, f, tau):
    """N(t) -> S(t+1), purely local."""
    n = s.F[:, s.own].copy()
    fl = s.fluid[s.own]
    rho, jx, jy, jz = R.moments(n, *f)
    ns = R.collide(n, rho, jx, jy, jz, *f, fl, tau)
    for l in range(19):
        s.F[INV[l], s.own] = np.where(fl, ns[l], s.F[INV[l], s.own])


def odd_step(s, f, tau):
    """S(t+1) -> N(t+2): pull from slot inv(l) of r - c_l (bounce-back: slot l of r), collide, push into slot l
    of r + c_l (bounce-back: slot inv(l) of r).  Reads first, then writes: every slot has one owner thread."""
    fl = s.fluid[s.own]
    n = np.zeros((19,) + fl.shape)
    for l in range(19):
        src_fluid = s.shifted(s.fluid, -C[l])
        pulled = s.shifted(s.F[INV[l]], -C[l])
        n[l] = np.where(fl, np.where(src_fluid, pulled, s.F[l, s.own]), 0.0)
    rho, jx, jy, jz = R.moments(n, *f)
    ns = R.collide(n, rho, jx, jy, jz, *f, fl, tau)
    for l in range(19):
        dst_fluid = s.shifted(s.fluid, C[l])
        # bounce-back: own slot inv(l)
        s.F[INV[l], s.own] = np.where(fl & ~dst_fluid, ns[l], s.F[INV[l], s.own])
        # push: slot l of r + c_l (may be a halo plane)
        val = np.where(fl & dst_fluid, ns[l], np.nan)
        tgt = np.full(s.nat.shape, np.nan)
        tgt[1 + C[l][2]: 1 + C[l][2] + s.nzl] = val
        tgt = np.roll(tgt, shift=(C[l][1], C[l][0]), axis=(1, 2))
        s.F[l] = np.where(np.isnan(tgt), s.F[l], tgt)
    return n


def forward_exchange(slabs):
    R_ = len(slabs)
    for r, s in enumerate(slabs):
        up, dn = slabs[(r + 1) % R_], slabs[(r - 1) % R_]
        for m in DOWN:
            up.F[m, 0] = s.F[m, s.nzl]               # my top plane -> upper neighbour's lower halo
        for m in UP:
            dn.F[m, dn.nzl + 1] = s.F[m, 1]          # my bottom plane -> lower neighbour's upper halo


def reverse_exchange(slabs, masked=True):
    R_ = len(slabs)
    for r, s in enumerate(slabs):
        up, dn = slabs[(r + 1) % R_], slabs[(r - 1) % R_]
        for m in UP:                                 # my upper halo -> upper neighbour's bottom own plane
            owner = np.roll(s.fluid[s.nzl], shift=(C[m][1], C[m][0]), axis=(0, 1))   # fluid(h - c_m), h in the halo plane
            take = owner if masked else np.ones_like(owner)
            up.F[m, 1] = np.where(take & up.fluid[1], s.F[m, s.nzl + 1], up.F[m, 1])
        for m in DOWN:                               # my lower halo -> lower neighbour's top own plane
            owner = np.roll(s.fluid[1], shift=(C[m][1], C[m][0]), axis=(0, 1))
            take = owner if masked else np.ones_like(owner)
            dn.F[m, dn.nzl] = np.where(take & dn.fluid[dn.nzl], s.F[m, 0], dn.F[m, dn.nzl])


def poison_halos(slabs, keep_lower, keep_upper):
    for s in slabs:
        for m in range(19):
            if m not in keep_lower:
                s.F[m, 0] = np.nan
            if m not in keep_upper:
                s.F[m, s.nzl + 1] = np.nan


def run(nat, nslabs, steps, tau=0.9, masked=True):
    lz = nat.shape[0]
    fluid = nat == 0
    f = (np.float64(1e-3), np.float64(-2e-3), np.float64(5e-4))
    # single-domain reference (closed-form pull rule), N layout after every step
    n = np.stack([np.where(fluid, R.W[l], 0.0) for l in range(19)])
    rho, jx, jy, jz = R.moments(n, *f)
    bounds = [(lz * r) // nslabs for r in range(nslabs + 1)]
    slabs = [Slab(nat, bounds[r], bounds[r + 1] - bounds[r]) for r in range(nslabs)]
    for s in slabs:
        s.F[:, s.own] = n[:, s.k0:s.k0 + s.nzl]
    worst = 0
    for t in range(steps):
        ns = R.collide(n, rho, jx, jy, jz, *f, fluid, tau)
        n = R.pull_closed_form(ns, nat)
        rho, jx, jy, jz = R.moments(n, *f)
        if t % 2 == 0:
            for s in slabs:
                even_step(s, f, tau)
            forward_exchange(slabs)
            poison_halos(slabs, keep_lower=DOWN, keep_upper=UP)
            # S(t+1): slot inv(l) holds n*(t+1)(r,l)
            for s in slabs:
                for l in range(19):
                    got = s.F[INV[l], s.own]
                    worst += int((got[s.fluid[s.own]] != ns[l][s.k0:s.k0 + s.nzl][s.fluid[s.own]]).sum())
        else:
            for s in slabs:
                odd_step(s, f, tau)
            reverse_exchange(slabs, masked=masked)
            for s in slabs:
                fl = s.fluid[s.own]
                for l in range(19):
                    worst += int((s.F[l, s.own][fl] != n[l][s.k0:s.k0 + s.nzl][fl]).sum())
    return worst


@pytest.mark.parametrize("shape,nslabs,p,seed", [((8, 5, 6), 2, 0.3, 1), ((9, 4, 5), 3, 0.25, 2), ((4, 6, 7), 4, 0.35, 3),
                                                 ((6, 1, 9), 2, 0.3, 4), ((5, 7, 1), 5, 0.2, 5)])
def test_masked_protocol_is_exact(shape, nslabs, p, seed):
    lz, ly, lx = shape
    nat = random_nature(lx, ly, lz, p, seed)
    assert run(nat, nslabs, steps=6) == 0


def test_plain_return_copy_corrupts_bounce_back_slots():
    """Without the owner mask the stale forward copy overwrites the receiver's bounce-back values."""
    nat = random_nature(6, 5, 8, 0.3, 1)
    assert run(nat, 2, steps=4, masked=False) > 0
