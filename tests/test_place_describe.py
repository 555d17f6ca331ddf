"""CPU: the host side of rbgtopo_place_groups' DIRECT path (rbgtopo_place_describe runs the two validation passes and
the launch-order computation of the call itself, without a device): totals and the launch geometry agree with the
staged path's plan geometry (rbgtopo_plan_describe), the launch order is the stable descending-weight permutation,
and malformed GROUPS blobs get the codes the staged path gives them."""
import ctypes as C

import numpy as np
import pytest

from rbg_b200 import _lib, synth
from rbg_b200.engine import plan_steps
from rbg_b200.plugin import EXCLUSIVE_TOPOLOGY_KEY, GANG_SCHEDULING_KEY, B200TopoPodGroupManager, RoleBasedGroup, RoleSpec
from test_plugin_host import OraclePlacer

I32P = C.POINTER(C.c_int32)
HDR, GW = 8, 12


def fleet(n_nodes, n_groups, seed, excl_every=0, gang_every=0, big_every=0):
    shapes = [synth.shape_mooncake(), synth.shape_pd_144(), synth.shape_fleet8(), synth.shape_sglang_pd()]
    out = []
    for g in range(n_groups):
        sh = shapes[g % len(shapes)]
        roles = [RoleSpec(r.name, r.replicas, tuple(r.deps), r.demand) for r in sh.roles]
        if big_every and g % big_every == 0:
            roles[1].replicas = 41
        ann = {}
        if excl_every and g % excl_every == 0:
            ann[EXCLUSIVE_TOPOLOGY_KEY] = "topology.kubernetes.io/nvlink-domain"
        if gang_every and g % gang_every == 0:
            ann[GANG_SCHEDULING_KEY] = "true"
        placed = [(sh.roles[q].name, node) for node, q, _ in synth.random_anchors(n_nodes, len(sh.roles), g % 3, seed, g)]
        out.append(RoleBasedGroup("default", f"rbg{g}", roles, annotations=ann, gid=g, policy_rules=sh.policy_rules, placed=placed))
    return out


def describe(gb, topo, degp1, wsum):
    lib = _lib.load()
    gb = np.ascontiguousarray(gb, dtype=np.int32)
    order = np.full(max(1, int(gb[2])), -1, dtype=np.int32)
    geom = np.zeros(8, dtype=np.int32)
    rc = lib.rbgtopo_place_describe(gb.ctypes.data_as(I32P), len(gb), topo.n, len(topo.domain_owner), degp1.ctypes.data_as(I32P), wsum,
                                    order.ctypes.data_as(I32P), len(order), geom.ctypes.data_as(I32P))
    return rc, geom, order


def plan_rc(gb, topo, degp1, wsum):
    lib = _lib.load()
    gb = np.ascontiguousarray(gb, dtype=np.int32)
    ns, nw, pw = C.c_int32(), C.c_int32(), C.c_int64()
    return lib.rbgtopo_plan_describe(gb.ctypes.data_as(I32P), len(gb), topo.n, len(topo.domain_owner), degp1.ctypes.data_as(I32P), wsum,
                                     None, 0, C.byref(ns), C.byref(nw), C.byref(pw))


@pytest.fixture(scope="module")
def world():
    n = 2048
    topo = synth.make_topology(n, seed=6, tiers=3)
    degp1 = (np.diff(topo.row_ptr) + 1).astype(np.int32)
    wsum = int(max(topo.edge_w[topo.row_ptr[i]:topo.row_ptr[i + 1]].sum() for i in range(n)))
    return topo, degp1, wsum


@pytest.mark.parametrize("kw", [dict(n_groups=12, seed=2, excl_every=4, gang_every=5), dict(n_groups=64, seed=9, big_every=5),
                                dict(n_groups=1, seed=1), dict(n_groups=33, seed=4, excl_every=2)])
def test_geometry_agrees_with_the_staged_plan(world, kw):
    topo, degp1, wsum = world
    gb, _ = B200TopoPodGroupManager(OraclePlacer(topo)).groups_blob(fleet(topo.n, **kw))
    gb = np.ascontiguousarray(gb, dtype=np.int32)
    rc, geom, order = describe(gb, topo, degp1, wsum)
    assert rc == 0
    steps = plan_steps(gb, topo.n, len(topo.domain_owner))          # (group, wave, sec, end, rep, row, next, i0) per step
    ng = int(gb[2])
    pend = [int(gb[HDR + g * GW + 9]) for g in range(ng)]
    na = [int(gb[HDR + g * GW + 6]) for g in range(ng)]
    assert geom[0] == int(gb[4]) == sum(pend)
    active = [g for g in range(ng) if pend[g] > 0]
    assert geom[1] == len(active) == sum(1 for st in steps if st[1] == 0)
    assert geom[2] == max(int(gb[HDR + g * GW + 3]) for g in active)
    # per step: role rows = next step's row offset - this one's (wave-major); the last step closes with the total
    rows = sorted(int(st[5]) for st in steps)
    # the largest table: last wave of a group = replicas placed before it (i0) + neighbourhoods of scheduled pods and of those replicas
    maxdeg = int(degp1.max())
    want_cap = 0
    for g in active:
        last = max((st for st in steps if st[0] == g), key=lambda st: st[1])
        anc = int(gb[HDR + g * GW + 7])
        pcp = sum(int(degp1[int(gb[anc + 3 * a])]) for a in range(na[g]))
        want_cap = max(want_cap, int(last[7]) + pcp + int(last[7]) * maxdeg)
    assert geom[4] == want_cap
    assert geom[7] >= geom[4] and geom[7] % 32 == 0 and geom[6] > geom[7] and (geom[6] & (geom[6] - 1)) == 0
    assert geom[5] == max(128, 32 * geom[3]) and 1 <= geom[3] <= 8 and len(rows) == len(steps)
    # launch order: stable, descending (pending replicas + scheduled pods)
    got = [int(x) for x in order[:geom[1]]]
    assert sorted(got) == active
    wmax = max(pend[g] + na[g] for g in active)
    key = lambda g: 1023 - (pend[g] + na[g]) * 1023 // wmax          # the 1 024 weight buckets of the counting sort
    assert got == sorted(active, key=lambda g: (key(g), g))


def test_bad_blobs_get_the_staged_paths_codes(world):
    topo, degp1, wsum = world
    gb, _ = B200TopoPodGroupManager(OraclePlacer(topo)).groups_blob(fleet(topo.n, 12, seed=2, excl_every=4, gang_every=5))
    gb = np.ascontiguousarray(gb, dtype=np.int32)
    assert describe(gb, topo, degp1, wsum)[0] == 0 and plan_rc(gb, topo, degp1, wsum) == 0
    g3 = HDR + 3 * GW
    roles3, pair3 = int(gb[g3 + 4]), int(gb[g3 + 5])
    g_anc = next(g for g in range(12) if int(gb[HDR + g * GW + 6]) > 0)
    ra = HDR + g_anc * GW
    a_off, q_anc, pair_anc = int(gb[ra + 7]), int(gb[ra + 3]), int(gb[ra + 5])
    role_anc = int(gb[a_off + 1])
    cases = [(0, 0x12345, -1), (g3 + 3, 0, -6), (g3 + 3, 17, -6), (roles3 + 1, -2, -1), (roles3 + 2, 1 << 20, -1), (roles3 + 3, 0x40, -1),
             (pair3, -1, -1), (g3 + 1, 0x100, -1), (g3 + 8, int(gb[g3 + 8]) + 1, -1), (g3 + 2, 1 << 20, -1), (4, int(gb[4]) + 1, -1),
             (pair_anc + role_anc, 1 << 23, -4), (a_off, topo.n + 5, -1), (a_off + 1, 99, -1), (a_off + 2, (1 << 24) + 1, -1),
             (g3 + 4, len(gb), -1), (g3 + 7, -5, -1)]
    for idx, val, code in cases:
        bad = gb.copy()
        bad[idx] = val
        assert describe(bad, topo, degp1, wsum)[0] == code, (idx, val)
        assert plan_rc(bad, topo, degp1, wsum) == code, (idx, val)          # the staged path's validation agrees


def test_shape_caches_do_not_leak_between_blobs(world):
    """The per-shape caches compare CONTENT: the same role table with another pair matrix, or another blob altogether,
    is validated afresh (a stale hit would accept the negative weight)."""
    topo, degp1, wsum = world
    mk = lambda seed: np.ascontiguousarray(B200TopoPodGroupManager(OraclePlacer(topo)).groups_blob(fleet(topo.n, 8, seed=seed))[0], dtype=np.int32)
    a = mk(2)
    assert describe(a, topo, degp1, wsum)[0] == 0
    b = a.copy()
    b[int(b[HDR + 5 * GW + 5])] = -7                                 # group 5's pair matrix, same role table as group 1 (shapes repeat every 4)
    assert describe(b, topo, degp1, wsum)[0] == -1
    assert describe(a, topo, degp1, wsum)[0] == 0
    for _ in range(3):                                               # alternate: hits and misses interleave
        assert describe(mk(3), topo, degp1, wsum)[0] == 0
        assert describe(b, topo, degp1, wsum)[0] == -1


def test_the_parallel_host_loops_in_a_subprocess():
    """From RBGTOPO_HOST_PARALLEL_MIN groups on (default 4 096) the validation loops run under OpenMP with per-thread shape
    caches; the switch is read when the library loads, hence the subprocess: the same tests with every loop parallel."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RBGTOPO_HOST_PARALLEL_MIN="1", RBGTOPO_HOST_THREADS="4")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_place_describe.py", "tests/test_plan_describe.py", "-q", "-x",
                        "-k", "not subprocess"], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "passed" in r.stdout
