"""CPU: property-based checks of the C oracle on small random instances (hypothesis) — the
second restatement agrees bit for bit, and the size-independent properties the domain offers
hold: capacity is never exceeded, gang steps are all-or-nothing, an exclusive step stays in
one domain that no other group owns, scores are linear in the anchor counts, the lists are
strictly descending."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import placer, placer_ref
from rbg_b200 import synth
from rbg_b200.blob import ROLE_EXCLUSIVE, STEP_EXCLUSIVE, STEP_GANG, BlobBuilder, Step

CFG = dict(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.too_slow])


@st.composite
def instances(draw):
    n = draw(st.integers(1, 48))
    tiers = draw(st.integers(1, 3))
    topo = synth.make_topology(n, seed=draw(st.integers(0, 10 ** 6)), tiers=tiers,
                               owned_frac=draw(st.sampled_from([0.0, 0.3])), max_free=draw(st.integers(1, 6)))
    if draw(st.booleans()):                       # scarce capacity
        topo.free[draw(st.integers(0, n - 1))::2] = 0
    steps = []
    for s in range(draw(st.integers(1, 4))):
        P = draw(st.integers(1, 4))
        Q = P + draw(st.integers(0, 2))
        roles, left = [], 32
        for _ in range(P):
            cnt = draw(st.integers(1, min(8, left - (P - len(roles) - 1))))
            left -= cnt
            roles.append((cnt, draw(st.integers(0, 3)), draw(st.integers(0, 16)),
                          ROLE_EXCLUSIVE if draw(st.booleans()) else 0))
        pair = [[draw(st.integers(0, 2)) for _ in range(Q)] for _ in range(P)]
        anc = [(draw(st.integers(0, n - 1)), draw(st.integers(0, Q - 1)), draw(st.integers(0, 2)))
               for _ in range(draw(st.integers(0, 4)))]
        cons = [(draw(st.integers(0, n - 1)), draw(st.integers(1, 3))) for _ in range(draw(st.integers(0, 3)))]
        excl = draw(st.booleans())
        flags = (STEP_EXCLUSIVE if excl else 0) | (STEP_GANG if draw(st.booleans()) else 0)
        fixed = draw(st.integers(0, len(topo.domain_owner) - 1)) if (excl and draw(st.booleans())) else -1
        steps.append(Step(gid=s, roles=roles, pair=pair, anchors=anc, consumed=cons, flags=flags, fixed_domain=fixed))
    return topo, steps


def _blob(steps):
    bb = BlobBuilder()
    for s in steps:
        bb.add(s)
    return bb.build()


@settings(**CFG)
@given(instances())
def test_restatements_agree_and_invariants_hold(inst):
    topo, steps = inst
    blob = _blob(steps)
    a = placer.place(topo, blob)
    assert a["rc"] == 0
    b = placer_ref.place(topo, blob)
    assert np.array_equal(a["matrix"].view(np.uint32), b["matrix"].view(np.uint32))
    assert np.array_equal(a["topk"], b["topk"]) and np.array_equal(a["assign"], b["assign"])
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["domain"], b["domain"])
    rep = rr = 0
    for si, s in enumerate(steps):
        R = s.n_replicas
        nodes = a["assign"][rep:rep + R]
        # capacity: demand placed on a node never exceeds free - consumed
        used = {}
        k = 0
        for cnt, demand, _, _ in s.roles:
            for _ in range(cnt):
                if nodes[k] >= 0:
                    used[int(nodes[k])] = used.get(int(nodes[k]), 0) + demand
                k += 1
        cons = {}
        for nd, amt in s.consumed:
            cons[nd] = cons.get(nd, 0) + amt
        for nd, u in used.items():
            assert u <= topo.free[nd] - cons.get(nd, 0), (si, nd)
        placed = nodes >= 0
        if s.flags & STEP_GANG:
            assert placed.all() or not placed.any()
            assert a["status"][si] in (0, 2) and (a["status"][si] == 0) == bool(placed.all())
        else:
            assert a["status"][si] == (0 if placed.all() else 1)
        if s.flags & STEP_EXCLUSIVE:
            k = 0
            doms = set()
            for cnt, _, _, rflags in s.roles:
                for _ in range(cnt):
                    if nodes[k] >= 0 and (rflags & ROLE_EXCLUSIVE):
                        doms.add(int(topo.domain[nodes[k]]))
                    k += 1
            assert len(doms) <= 1
            for d in doms:
                assert d == a["domain"][si] and topo.domain_owner[d] in (-1, s.gid)
                assert s.fixed_domain in (-1, d)
        # every list is strictly descending and zero-padded
        for p in range(len(s.roles)):
            keys = a["topk"][rr + p]
            nz = keys[keys != 0]
            assert all(int(x) > int(y) for x, y in zip(nz, nz[1:])) and (keys[len(nz):] == 0).all()
        rep += R
        rr += len(s.roles)


@settings(**CFG)
@given(instances(), st.integers(2, 3))
def test_scores_are_linear_in_the_anchor_counts(inst, factor):
    """S = W·A is linear: with need = 0, multiplying every anchor count by `factor` multiplies every
    finite score by `factor` (the exactness contract permitting) and leaves feasibility untouched."""
    topo, steps = inst
    base, scaled = [], []
    for s in steps:
        roles = [(c, d, 0, f) for c, d, _, f in s.roles]
        base.append(Step(s.gid, roles, s.pair, s.anchors, s.consumed, s.flags & ~STEP_EXCLUSIVE))
        scaled.append(Step(s.gid, roles, s.pair, [(n_, q, c * factor) for n_, q, c in s.anchors], s.consumed,
                           s.flags & ~STEP_EXCLUSIVE))
    a, b = placer.place(topo, _blob(base)), placer.place(topo, _blob(scaled))
    if a["rc"] != 0 or b["rc"] != 0:
        assert -4 in (a["rc"], b["rc"])           # only the exactness bound may refuse
        return
    fa, fb = np.isfinite(a["matrix"]), np.isfinite(b["matrix"])
    assert np.array_equal(fa, fb)
    assert np.array_equal(a["matrix"][fa] * factor, b["matrix"][fb])
