"""world_size = 2 on CPU (gloo): the node-axis sharding protocol of SURVEY.md §8e —
every rank selects a local top-K per role row on its slab, ONE all-gather of the
key lists, identical merge + greedy on every rank; exclusive steps without a
fixed domain take a second all-gather restricted to D*.  The per-rank scoring is
played by the CPU oracle's dense matrix (test infrastructure); what is under test
is the protocol the CUDA shard calls implement (rbgtopo_shard_score / _merge /
_assign) and the torch.distributed plumbing around it."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import placer as oracle_placer
from oracle.placer_ref import _key
from rbg_b200 import synth
from rbg_b200.blob import ROLE_EXCLUSIVE, STEP_EXCLUSIVE, STEP_GANG, BlobBuilder, Step

KS = 32


def slab_bounds(n, world):
    """compute_slab() of rbg_b200/csrc/rbgtopo.cu: boundaries rounded down to 128."""
    def bound(g):
        if g <= 0:
            return 0
        if g >= world:
            return n
        return (g * n // world) // 128 * 128
    return [(bound(g), bound(g + 1)) for g in range(world)]


def local_lists(matrix, topo, blob, lo, hi, dstar=None):
    """Rank-local top-K_p keys per role row; pass 1 (dstar None) selects exclusive
    roles of unknown-domain steps unrestricted, pass 2 restricts them to dstar[s]."""
    ns, tp = int(blob[2]), int(blob[5])
    out = np.zeros((tp, KS), dtype=np.uint64)
    for s in range(ns):
        st = blob[8 + 16 * s: 8 + 16 * (s + 1)]
        flags, fixed, P, role_off, rep_off, rr_off = int(st[1]), int(st[2]), int(st[3]), int(st[4]), int(st[12]), int(st[13])
        roles = blob[role_off: role_off + 4 * P].reshape(P, 4)
        unknown = bool(flags & STEP_EXCLUSIVE) and fixed < 0
        kacc, row = 0, rep_off
        for p in range(P):
            kacc += int(roles[p, 0])
            rexcl = bool(flags & STEP_EXCLUSIVE) and bool(roles[p, 3] & ROLE_EXCLUSIVE)
            dom = None
            if rexcl:
                dom = fixed if not unknown else (None if dstar is None else int(dstar[s]))
            if dstar is not None and not (unknown and rexcl):
                row += int(roles[p, 0])
                continue
            vals = matrix[row, lo:hi]
            idx = np.nonzero(np.isfinite(vals))[0] + lo
            if dom is not None:
                idx = idx[topo.domain[idx] == dom] if dom >= 0 else idx[:0]
            keys = sorted((_key(int(matrix[row, i]), int(i)) for i in idx), reverse=True)[:min(kacc, topo.n)]
            out[rr_off + p, :len(keys)] = np.array(keys, dtype=np.uint64)
            row += int(roles[p, 0])
    return out


def merge_and_greedy(topo, blob, lists_all, excl_all):
    ns = int(blob[2])
    assign = np.full(int(blob[4]), -1, dtype=np.int32)
    status = np.zeros(ns, dtype=np.int32)
    dstar = np.full(ns, -1, dtype=np.int32)
    merged = {}
    for s in range(ns):
        st = blob[8 + 16 * s: 8 + 16 * (s + 1)]
        flags, fixed, P, role_off, rr_off = int(st[1]), int(st[2]), int(st[3]), int(st[4]), int(st[13])
        roles = blob[role_off: role_off + 4 * P].reshape(P, 4)
        kacc, dset = 0, not (flags & STEP_EXCLUSIVE) or fixed >= 0
        if flags & STEP_EXCLUSIVE and fixed >= 0:
            dstar[s] = fixed
        for p in range(P):
            kacc += int(roles[p, 0])
            keys = sorted((int(k) for part in lists_all for k in part[rr_off + p] if k), reverse=True)[:kacc]
            merged[(s, p)] = keys
            if not dset and roles[p, 3] & ROLE_EXCLUSIVE:
                dstar[s] = topo.domain[0xFFFFFFFF - (keys[0] & 0xFFFFFFFF)] if keys else -1
                dset = True
    if excl_all is None:
        return merged, dstar
    for s in range(ns):
        st = blob[8 + 16 * s: 8 + 16 * (s + 1)]
        flags, fixed, P, role_off, ncons, cons_off, R, rep_off, rr_off = (int(st[i]) for i in (1, 2, 3, 4, 9, 10, 11, 12, 13))
        roles = blob[role_off: role_off + 4 * P].reshape(P, 4)
        avail = topo.free.astype(np.int64).copy()
        for c in blob[cons_off: cons_off + 2 * ncons].reshape(ncons, 2):
            avail[c[0]] -= c[1]
        kacc, r, unplaced = 0, rep_off, 0
        for p in range(P):
            kacc += int(roles[p, 0])
            keys = merged[(s, p)]
            if flags & STEP_EXCLUSIVE and fixed < 0 and roles[p, 3] & ROLE_EXCLUSIVE:
                keys = sorted((int(k) for part in excl_all for k in part[rr_off + p] if k), reverse=True)[:kacc]
            for _ in range(int(roles[p, 0])):
                pick = -1
                for k in keys:
                    node = 0xFFFFFFFF - (k & 0xFFFFFFFF)
                    if avail[node] >= roles[p, 1]:
                        pick = node
                        break
                if pick >= 0:
                    avail[pick] -= roles[p, 1]
                else:
                    unplaced += 1
                assign[r] = pick
                r += 1
        if unplaced and flags & STEP_GANG:
            assign[rep_off: rep_off + R] = -1
            status[s] = 2
        else:
            status[s] = 1 if unplaced else 0
    return assign, status, dstar


def _worker(rank, world, port, n, seed):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        topo = synth.make_topology(n, seed=seed, tiers=3, owned_frac=0.25)
        rng = np.random.default_rng(seed)
        bb = BlobBuilder()
        for s in range(10):
            P = int(rng.integers(1, 4))
            roles = [(int(rng.integers(1, 5)), int(rng.integers(0, 3)), int(rng.integers(0, 9)),
                      ROLE_EXCLUSIVE if rng.random() < 0.8 else 0) for _ in range(P)]
            anc = [(int(rng.integers(0, n)), int(rng.integers(0, P)), 1) for _ in range(int(rng.integers(0, 4)))]
            excl = s % 2 == 1
            bb.add(Step(gid=s, roles=roles, pair=rng.integers(0, 3, size=(P, P)).tolist(), anchors=anc,
                        consumed=[(int(rng.integers(0, n)), 1)] if s % 3 == 0 else [],
                        flags=(STEP_EXCLUSIVE if excl else 0) | (STEP_GANG if s % 4 == 0 else 0),
                        fixed_domain=int(rng.integers(0, len(topo.domain_owner))) if (excl and s % 4 == 1) else -1))
        blob = bb.build()
        ref = oracle_placer.place(topo, blob)
        assert ref["rc"] == 0
        lo, hi = slab_bounds(n, world)[rank]
        mine = local_lists(ref["matrix"], topo, blob, lo, hi)
        parts = [torch.zeros_like(torch.from_numpy(mine.view(np.int64))) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(mine.view(np.int64)))                     # collective 1
        lists_all = [p.numpy().view(np.uint64) for p in parts]
        _, dstar = merge_and_greedy(topo, blob, lists_all, None)
        mine2 = local_lists(ref["matrix"], topo, blob, lo, hi, dstar=dstar)
        parts2 = [torch.zeros_like(torch.from_numpy(mine2.view(np.int64))) for _ in range(world)]
        dist.all_gather(parts2, torch.from_numpy(mine2.view(np.int64)))                   # collective 2 (exclusive)
        assign, status, dstar = merge_and_greedy(topo, blob, lists_all, [p.numpy().view(np.uint64) for p in parts2])
        assert np.array_equal(assign, ref["assign"]), (rank, assign, ref["assign"])
        assert np.array_equal(status, ref["status"])
        assert np.array_equal(dstar, ref["domain"])
        # every rank must hold the identical result
        box = [None] * world
        dist.all_gather_object(box, assign.tolist())
        assert all(b == box[0] for b in box)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,seed", [(1000, 1), (2500, 2)])
def test_sharded_protocol_world2_gloo(n, seed):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n, seed), nprocs=2, join=True)


def test_slab_bounds_cover_the_node_axis():
    for n in (1, 127, 128, 1000, 10000, 50000):
        for w in (1, 2, 4, 8):
            b = slab_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert all(lo % 128 == 0 for lo, _ in b)
