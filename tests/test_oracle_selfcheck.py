"""The C oracle against the independent numpy restatement of the same spec, and
against hand-computed known answers.  (Parity for this path is unpinned upstream:
the reference has no scoring code, SURVEY.md §0 — these are the pins we can have.)"""
import numpy as np

from oracle import placer, placer_ref
from rbg_b200 import synth
from rbg_b200.blob import ROLE_EXCLUSIVE, STEP_EXCLUSIVE, STEP_GANG, BlobBuilder, Step


def _rand_blob(topo, seed, n_steps, excl):
    rng = np.random.default_rng(seed)
    bb = BlobBuilder()
    for s in range(n_steps):
        P = int(rng.integers(1, 5))
        Q = P + int(rng.integers(0, 2))
        roles = [(int(rng.integers(1, 5)), int(rng.integers(0, 4)), int(rng.integers(0, 17)),
                  ROLE_EXCLUSIVE if rng.random() < 0.8 else 0) for _ in range(P)]
        pair = rng.integers(0, 3, size=(P, Q)).tolist()
        anc = [(int(rng.integers(0, topo.n)), int(rng.integers(0, Q)), int(rng.integers(1, 3)))
               for _ in range(int(rng.integers(0, 5)))]
        cons = [(int(rng.integers(0, topo.n)), int(rng.integers(1, 4))) for _ in range(int(rng.integers(0, 4)))]
        flags = (STEP_EXCLUSIVE if excl else 0) | (STEP_GANG if rng.random() < 0.3 else 0)
        fixed = int(rng.integers(0, len(topo.domain_owner))) if (excl and rng.random() < 0.4) else -1
        bb.add(Step(gid=s, roles=roles, pair=pair, anchors=anc, consumed=cons, flags=flags, fixed_domain=fixed))
    return bb.build()


def test_c_oracle_matches_numpy_restatement():
    for n, tiers, owned in [(5, 1, 0.0), (64, 2, 0.0), (300, 3, 0.3), (1000, 4, 0.2)]:
        topo = synth.make_topology(n, seed=n, tiers=tiers, owned_frac=owned)
        assert placer.check_topology(topo) == 0
        for seed in range(4):
            blob = _rand_blob(topo, 1000 * n + seed, 6, excl=bool(seed & 1))
            a = placer.place(topo, blob)
            b = placer_ref.place(topo, blob)
            assert a["rc"] == 0
            assert np.array_equal(a["matrix"].view(np.uint32), b["matrix"].view(np.uint32))
            assert np.array_equal(a["topk"], b["topk"])
            assert np.array_equal(a["assign"], b["assign"])
            assert np.array_equal(a["status"], b["status"])
            assert np.array_equal(a["domain"], b["domain"])


def test_known_answer_path_graph():
    """4 nodes in a path 0-1-2-3, weights 1000/100/10; hand-computed scores."""
    topo = synth.Topology(row_ptr=np.array([0, 1, 3, 5, 6], np.int32), col_idx=np.array([1, 0, 2, 1, 3, 2], np.int32),
                          edge_w=np.array([1000, 1000, 100, 100, 10, 10], np.int32),
                          free=np.array([1, 2, 9, 0], np.int32), domain=np.array([0, 0, 1, 1], np.int32),
                          domain_owner=np.array([-1, -1], np.int32))
    # one role, need = 1, demand 1, no anchors: A = min(free, 8) = [1, 2, 8, 0]
    # S0 = 1000*2 + 8000*1 = 10000; S1 = 1000*1 + 100*8 + 8000*2 = 17800
    # S2 = 100*2 + 10*0 + 8000*8 = 64200; S3 = 10*8 + 0 = 80 but free 0 < demand -> -inf
    blob = BlobBuilder().add(Step(gid=0, roles=[(2, 1, 1, 1)], pair=[[0]])).build()
    r = placer.place(topo, blob)
    assert r["matrix"][0].tolist() == [10000.0, 17800.0, 64200.0, -np.inf]
    assert r["assign"].tolist() == [2, 2]          # node 2 has 9 free slots
    # with an anchor pod of role 0 on node 0 and pair = 3: A += 3 at node 0
    # S0 += 8000*3 = 34000; S1 += 1000*3 = 20800; others unchanged
    blob = BlobBuilder().add(Step(gid=0, roles=[(1, 1, 1, 1)], pair=[[3]], anchors=[(0, 0, 1)])).build()
    r = placer.place(topo, blob)
    assert r["matrix"][0].tolist() == [34000.0, 20800.0, 64200.0, -np.inf]
    # consumed capacity makes node 2 infeasible for demand 9
    blob = BlobBuilder().add(Step(gid=0, roles=[(1, 9, 1, 1)], pair=[[0]], consumed=[(2, 1)])).build()
    r = placer.place(topo, blob)
    assert r["assign"].tolist() == [-1] and r["status"].tolist() == [1]


def test_oracle_threads_agree():
    topo = synth.make_topology(2000, seed=9, tiers=4)
    blob = _rand_blob(topo, 77, 40, excl=False)
    a = placer.place(topo, blob, nthreads=1)
    b = placer.place(topo, blob, nthreads=max(2, placer.max_threads()))
    assert np.array_equal(a["assign"], b["assign"]) and np.array_equal(a["matrix"].view(np.uint32), b["matrix"].view(np.uint32))


def test_exactness_violation_is_reported():
    topo = synth.make_topology(64, seed=1, tiers=2)
    blob = BlobBuilder().add(Step(gid=0, roles=[(1, 0, 16, 1)], pair=[[500]], anchors=[(3, 0, 50)])).build()
    assert placer.place(topo, blob)["rc"] == -4
