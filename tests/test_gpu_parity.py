"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU
oracle on the same seeded inputs — bit-exact dense matrix, top-K keys, assignment.
NOTE: parity is against OUR frozen spec; the reference has no such path
(SURVEY.md §0), so this parity is "unpinned upstream"."""
import numpy as np
import pytest

from rbg_b200 import synth
from rbg_b200.blob import ROLE_EXCLUSIVE, STEP_EXCLUSIVE, STEP_GANG, BlobBuilder, Step

pytestmark = pytest.mark.gpu


def _pair(p, q, v=1):
    return [[v] * q for _ in range(p)]


def _random_steps(topo, seed, n_steps, excl=False, gang=False, anchors=6, max_roles=5):
    rng = np.random.default_rng(seed)
    bb = BlobBuilder()
    for s in range(n_steps):
        P = int(rng.integers(1, max_roles + 1))
        Q = P + int(rng.integers(0, 3))
        roles, left = [], 32
        for p in range(P):
            cnt = int(rng.integers(1, min(8, left - (P - p - 1)) + 1))
            left -= cnt
            flags = ROLE_EXCLUSIVE if (not excl or rng.random() < 0.8) else 0
            roles.append((cnt, int(rng.integers(0, 4)), int(rng.integers(0, 17)), flags))
        pair = rng.integers(0, 3, size=(P, Q)).tolist()
        na = int(rng.integers(0, anchors + 1))
        anc = [(int(rng.integers(0, topo.n)), int(rng.integers(0, Q)), int(rng.integers(1, 3))) for _ in range(na)]
        nc = int(rng.integers(0, 5))
        cons = [(int(rng.integers(0, topo.n)), int(rng.integers(1, 4))) for _ in range(nc)]
        if anc and rng.random() < 0.5:
            cons.append((anc[0][0], 1))
        flags = (STEP_EXCLUSIVE if excl else 0) | (STEP_GANG if (gang and rng.random() < 0.5) else 0)
        fixed = -1
        if excl and rng.random() < 0.4:
            fixed = int(rng.integers(0, len(topo.domain_owner)))
        bb.add(Step(gid=s, roles=roles, pair=pair, anchors=anc, consumed=cons, flags=flags, fixed_domain=fixed))
    return bb.build()


@pytest.mark.parametrize("n,tiers", [(4, 1), (37, 2), (1024, 2), (2048, 3), (2500, 4), (10000, 4)])
def test_parity_sizes(n, tiers):
    from gpu_util import check_batch, new_engine
    topo = synth.make_topology(n, seed=n, tiers=tiers)
    eng = new_engine(topo)
    for seed in range(3):
        check_batch(eng, topo, _random_steps(topo, 100 * n + seed, 12))
    eng.close()


def test_parity_100_seeds_cfg2():
    """SURVEY.md §7 minimum slice: cfg2 (3 roles x (1,4,4), 1 024 nodes), >= 100 seeds."""
    from gpu_util import check_batch, new_engine
    eng = None
    for seed in range(100):
        topo = synth.make_topology(1024, seed=seed, tiers=2)
        if eng is None:
            eng = new_engine(topo)
        else:
            eng.set_topology(topo.row_ptr, topo.col_idx, topo.edge_w, topo.free, topo.domain, topo.domain_owner)
        bb = BlobBuilder()
        anc = synth.random_anchors(topo.n, 3, seed % 5, seed, 0)
        bb.add(Step(gid=0, roles=[(4, 1, 9, 1), (4, 1, 9, 1)], pair=_pair(2, 3), anchors=anc))
        bb.add(Step(gid=0, roles=[(1, 0, 9, 1)], pair=_pair(1, 3), anchors=anc))
        check_batch(eng, topo, bb.build())
    eng.close()


def test_parity_exclusive_and_gang():
    from gpu_util import check_batch, new_engine
    topo = synth.make_topology(4096, seed=5, tiers=4, owned_frac=0.3)
    eng = new_engine(topo)
    for seed in range(6):
        check_batch(eng, topo, _random_steps(topo, 7000 + seed, 24, excl=True, gang=True))
    eng.close()


def test_parity_scarce_capacity():
    """Few feasible nodes: lists shorter than K, unplaced replicas, gang failures."""
    from gpu_util import check_batch, new_engine
    topo = synth.make_topology(512, seed=3, tiers=2, max_free=1)
    topo.free[5:] = 0
    eng = new_engine(topo)
    bb = BlobBuilder()
    bb.add(Step(gid=0, roles=[(6, 1, 4, 1)], pair=_pair(1, 1)))
    bb.add(Step(gid=1, roles=[(6, 1, 4, 1)], pair=_pair(1, 1), flags=STEP_GANG))
    bb.add(Step(gid=2, roles=[(2, 0, 4, 1), (3, 5, 2, 1)], pair=_pair(2, 2)))
    ref = check_batch(eng, topo, bb.build())
    assert ref["status"][1] == 2 and (ref["assign"][6:12] == -1).all()
    eng.close()


def test_empty_batch_and_errors():
    from gpu_util import new_engine
    from rbg_b200.engine import RbgTopoError
    topo = synth.make_topology(256, seed=1, tiers=2)
    eng = new_engine(topo)
    a, s, d = eng.score_assign(BlobBuilder().build())
    assert len(a) == 0 and len(s) == 0
    bad = BlobBuilder().add(Step(gid=0, roles=[(1, 0, 1, 1)], pair=_pair(1, 1), anchors=[(9999, 0, 1)])).build()
    with pytest.raises(RbgTopoError) as ei:
        eng.score_assign(bad)
    assert ei.value.code == -1
    huge = BlobBuilder().add(Step(gid=0, roles=[(1, 0, 16, 1)], pair=[[200]], anchors=[(3, 0, 100)])).build()
    with pytest.raises(RbgTopoError) as ei:
        eng.score_assign(huge)
    assert ei.value.code == -4   # RBGTOPO_EINEXACT
    # asymmetric CSR is rejected
    with pytest.raises(RbgTopoError):
        eng.set_topology([0, 1, 1], [1], [5], [1, 1], [0, 0], [-1])
    eng.close()


def test_update_nodes_changes_base():
    from gpu_util import check_batch, new_engine
    topo = synth.make_topology(2048, seed=11, tiers=3)
    eng = new_engine(topo)
    blob = _random_steps(topo, 42, 8)
    check_batch(eng, topo, blob)
    topo.free = ((topo.free.astype(np.int64) * 7 + 3) % 9).astype(np.int32)
    topo.domain_owner[::5] = 3
    eng.update_nodes(topo.free, topo.domain_owner, generation=2)
    check_batch(eng, topo, _random_steps(topo, 43, 8, excl=True))
    assert eng.stats()["generation"] == 2
    eng.close()


def test_full_size_properties_cfg3():
    """BASELINE full size (10 000 nodes, batched mooncake level-1 steps): the
    oracle checks a sample of steps bit-exactly; size-independent properties hold
    for all: identical steps give identical rows; capacity is never exceeded."""
    from gpu_util import new_engine
    from oracle import placer as oracle_placer
    topo = synth.make_topology(10000, seed=0, tiers=4)
    eng = new_engine(topo)
    bb = BlobBuilder()
    B = 256
    for g in range(B):
        bb.add(Step(gid=g, roles=[(1, 1, 5, 1), (3, 1, 5, 1), (1, 1, 5, 1)], pair=_pair(3, 5),
                    anchors=[((g * 37) % topo.n, 0, 1)], consumed=[((g * 37) % topo.n, 0)]))
    blob = bb.build()
    h = eng.stage(blob)
    eng.run_staged(h, 2)
    assign, status, domain = eng.fetch(h)
    ref = oracle_placer.place(topo, blob, want_matrix=False, want_topk=False, nthreads=oracle_placer.max_threads())
    assert np.array_equal(assign, ref["assign"]) and np.array_equal(status, ref["status"])
    # rows of replicas of one role are identical; steps with the same anchor match
    r0 = eng.read_scores(h, 1)
    r1 = eng.read_scores(h, 2)
    assert np.array_equal(r0.view(np.uint32), r1.view(np.uint32))
    # capacity: per step, demand placed on a node never exceeds free
    for g in range(0, B, 17):
        nodes = assign[g * 5:(g + 1) * 5]
        for nd in set(nodes.tolist()):
            if nd >= 0:
                assert (nodes == nd).sum() <= topo.free[nd]
    eng.release(h)
    eng.close()
