"""World-size-2 gloo test (CPU) of the slab host logic: partition, periodic ring, and the halo plan.

Each rank holds its slab of post-collision populations plus two halo planes, exchanges exactly the
planes lbg_halo_plan() names (5 populations up, 5 down) with torch.distributed send/recv, applies
the pull rule on its own planes, and must reproduce the global result bit for bit.  The pull rule
used as the checker is the numpy restatement (test infrastructure), not product code.
"""
import os
import socket

import numpy as np
import pytest

from tests.util import random_nature


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, q):
    import torch
    import torch.distributed as dist
    from laboetie_b200 import api, slab
    from oracle import numpy_restatement as R
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lx, ly, lz = shape
        nat = random_nature(lx, ly, lz, 0.3, 5)
        rng = np.random.default_rng(11)
        nstar = rng.random((19, lz, ly, lx)) * (nat == 0)
        ref = R.pull_closed_form(nstar, nat)
        k0, nzl = api.partition(lz, world, rank)
        below, above = slab.ring_neighbours(rank, world)
        up, down = api.halo_plan()
        nat_s = slab.slab_with_halo(nat, k0, nzl)
        loc = np.zeros((19, nzl + 2, ly, lx))
        loc[:, 1:-1] = nstar[:, k0:k0 + nzl]
        # exchange: top own plane of the up-going populations -> upper neighbour's lower halo, and vice versa
        for lst, src_plane, dst_plane, to, frm in ((up, nzl, 0, above, below), (down, 1, nzl + 1, below, above)):
            for l in lst:
                send = torch.from_numpy(np.ascontiguousarray(loc[l, src_plane]))
                recv = torch.empty_like(send)
                ops = [dist.P2POp(dist.isend, send, to), dist.P2POp(dist.irecv, recv, frm)]
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
                loc[l, dst_plane] = recv.numpy()
        # pull on the slab (z is not periodic inside a slab: the halos stand in), x/y periodic
        fluid = nat_s == 0
        out = np.zeros_like(loc)
        for l in range(19):
            c = R.C[l]
            shift = lambda a: np.roll(a, shift=(c[2], c[1], c[0]), axis=(0, 1, 2))  # value at r - c  # noqa: E731
            src_fluid = shift(fluid)
            out[l] = np.where(fluid, np.where(src_fluid, shift(loc[l]), loc[R.INV[l]]), 0.0)
        ok = np.array_equal(out[:, 1:-1], ref[:, k0:k0 + nzl])
        # only the populations named by the plan were needed: all others never read a halo plane
        q.put((rank, bool(ok), k0, nzl))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(6, 5, 9), (1, 4, 6), (7, 1, 2)])
def test_halo_plan_and_partition_with_gloo(shape):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shape, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(r[1] for r in res), res
    assert sorted(r[2] for r in res)[0] == 0 and sum(r[3] for r in res) == shape[2]
