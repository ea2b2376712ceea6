"""Host model of the two peer-to-peer protocols of the process-per-GPU path (laboetie_b200/csrc/api.cu).

No GPU: every rank is a small state machine that executes the same sequence of operations the library enqueues
(compute -> push + signal -> ... -> wait + unpack; contribute -> wait -> reduce), and a randomised scheduler
interleaves the ranks -- and delays every remote store -- in any order the hardware could.  Checked:

  * halo planes through the receive buffers (halo_exchange_p2p / wait_halo): with TWO parities per side, the data a
    rank unpacks for exchange e is always exchange e of the right neighbour, never e+2 written on top of it;
    with ONE buffer the same schedule search finds an overwrite (the model is sensitive);
  * scalar all-reduces through the mailboxes (p2p_allreduce_kernel): with two parities every rank reduces exactly
    the contributions of the same all-reduce, and all ranks obtain identical results.
"""
import random

import pytest


class Net:
    """Remote stores in flight: per (sender, receiver) FIFO (a copy-engine stream / a thread's stores are ordered),
    delivered whenever the scheduler says so."""

    def __init__(self, n):
        self.q = {(s, r): [] for s in range(n) for r in range(n)}

    def send(self, s, r, fn):
        self.q[(s, r)].append(fn)

    def deliverable(self):
        return [k for k, v in self.q.items() if v]

    def deliver(self, k):
        self.q[k].pop(0)()


def run_halo(nranks, nsteps, parities, seed):
    """Returns the number of corrupted unpacks.  Rank program per step e = 1..nsteps:
    push(e) to both neighbours (data then flag, ordered), then wait for both flags >= e, then unpack(e)."""
    rng = random.Random(seed)
    net = Net(nranks)
    stage = [[[None, None] for _ in range(parities)] for _ in range(nranks)]   # [rank][parity][side] = (sender, exchange)
    flags = [[0, 0] for _ in range(nranks)]                                     # [rank][side]: 0 = from lower, 1 = from upper
    pc = [("push", 1)] * nranks
    bad = 0
    done = 0
    while done < nranks:
        choices = [("rank", r) for r in range(nranks) if pc[r] is not None] + [("net", k) for k in net.deliverable()]
        kind, x = rng.choice(choices)
        if kind == "net":
            net.deliver(x)
            continue
        r = x
        op, e = pc[r]
        up, dn = (r + 1) % nranks, (r - 1) % nranks
        if op == "push":
            par = e % parities
            # my top plane -> upper neighbour's side 0; my bottom plane -> lower neighbour's side 1; flag after data
            for nb, side in ((up, 0), (dn, 1)):
                net.send(r, nb, lambda nb=nb, side=side, par=par, e=e, r=r: stage[nb][par].__setitem__(side, (r, e)))
                net.send(r, nb, lambda nb=nb, side=side, e=e: flags[nb].__setitem__(side, max(flags[nb][side], e)))
            pc[r] = ("wait", e)
        elif op == "wait":
            if flags[r][0] >= e and flags[r][1] >= e:      # p2p_wait_kernel
                par = e % parities
                if stage[r][par][0] != (dn, e) or stage[r][par][1] != (up, e):   # halo_unpack_kernel reads the slots
                    bad += 1
                pc[r] = ("push", e + 1) if e < nsteps else None
                if pc[r] is None:
                    done += 1
            # else: the wait kernel keeps spinning; the scheduler picks something else
    return bad


@pytest.mark.parametrize("nranks", [2, 3, 4, 8])
def test_two_parities_suffice_for_the_receive_buffers(nranks):
    for seed in range(60):
        assert run_halo(nranks, nsteps=12, parities=2, seed=seed) == 0


def test_one_receive_buffer_is_not_enough():
    """A rank that has received both halos of exchange e may push e+1 before its neighbour has unpacked e."""
    assert any(run_halo(nranks, 12, 1, seed) > 0 for nranks in (2, 3, 4) for seed in range(200))


def run_allreduce(nranks, ncalls, parities, seed):
    """Every rank contributes value (rank, call) to every mailbox, waits for all contributions of that call, then
    reduces.  Returns (mismatches, results differ between ranks)."""
    rng = random.Random(seed)
    net = Net(nranks)
    mail = [[[(0, None)] * nranks for _ in range(parities)] for _ in range(nranks)]   # [rank][parity][sender] = (seq, payload)
    pc = [("put", 1)] * nranks
    results = [[] for _ in range(nranks)]
    bad, done = 0, 0
    while done < nranks:
        choices = [("rank", r) for r in range(nranks) if pc[r] is not None] + [("net", k) for k in net.deliverable()]
        kind, x = rng.choice(choices)
        if kind == "net":
            net.deliver(x)
            continue
        r = x
        op, c = pc[r]
        par = c % parities
        if op == "put":
            for dst in range(nranks):   # payload then sequence number: one ordered store stream per (sender, receiver)
                net.send(r, dst, lambda dst=dst, par=par, c=c, r=r: mail[dst][par].__setitem__(r, (c, (r, c))))
            pc[r] = ("get", c)
        else:
            if all(mail[r][par][s][0] >= c for s in range(nranks)):
                vals = [mail[r][par][s][1] for s in range(nranks)]
                if any(v != (s, c) for s, v in enumerate(vals)):
                    bad += 1
                results[r].append(tuple(vals))
                pc[r] = ("put", c + 1) if c < ncalls else None
                if pc[r] is None:
                    done += 1
    differ = any(results[r] != results[0] for r in range(nranks))
    return bad, differ


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_two_parities_suffice_for_the_mailboxes(nranks):
    for seed in range(60):
        bad, differ = run_allreduce(nranks, ncalls=10, parities=2, seed=seed)
        assert bad == 0 and not differ


def test_one_mailbox_slot_is_not_enough():
    assert any(run_allreduce(n, 10, 1, seed)[0] > 0 for n in (2, 3) for seed in range(200))
