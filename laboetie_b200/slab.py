"""z-slab plumbing for one-process-per-GPU runs (host logic; torch.distributed is plumbing only).

The lattice is cut into contiguous z-slabs, the reference's own decomposition axis
(module_moment_propagation.f90:207, "we parallelize over k").  Rank r owns global planes
[k0, k0+nzl); its arrays carry one halo plane below and one above.  Per LB step the 5 populations
with cz=+1 leave through the top face and the 5 with cz=-1 through the bottom face
(lbg_halo_plan); the C library moves them with ncclSend/ncclRecv.  This module only slices
geometry, distributes the NCCL id and gathers results.
"""
import numpy as np

from . import api


def slab_with_halo(nature, k0, nzl):
    """Planes k0-1 .. k0+nzl of a global (lz, ly, lx) array, periodic in z (module_system.f90:99-112)."""
    lz = nature.shape[0]
    idx = np.arange(k0 - 1, k0 + nzl + 1) % lz
    return np.ascontiguousarray(nature[idx])


def make_slab_sim(nature_global, rank, nranks, device, unique_id=None):
    """Build this rank's handle from the global geometry (small lattices) and join the ring."""
    lz = nature_global.shape[0]
    k0, nzl = api.partition(lz, nranks, rank)
    sim = api.LaboetieGPU(slab_with_halo(nature_global, k0, nzl), device=device, lz_global=lz, k0=k0, slab=True)
    if nranks > 1:
        sim.comm_init(nranks, rank, unique_id)
    return sim


def broadcast_unique_id(dist, rank):
    """Rank 0 creates the NCCL id; everybody receives it through torch.distributed."""
    box = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def ring_neighbours(rank, nranks):
    return (rank - 1) % nranks, (rank + 1) % nranks   # (below, above)
