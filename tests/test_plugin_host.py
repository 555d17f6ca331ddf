"""Host-side plugin logic on CPU: RoleBasedGroup -> levels -> waves -> steps, the
feedback of placements into later waves, gang / exclusive carry-over.  The placer
behind the plugin here is the CPU oracle (test infrastructure only) — the GPU
tests run the same scenarios through the C ABI and compare."""
import json

import numpy as np
import pytest

from oracle import placer as oracle_placer
from rbg_b200 import synth
from rbg_b200.blob import STEP_EXCLUSIVE, STEP_GANG
from rbg_b200.plugin import (EXCLUSIVE_TOPOLOGY_KEY, GANG_SCHEDULING_KEY, PLACEMENT_HINT_KEY,
                             ROLE_DISABLE_EXCLUSIVE_KEY, B200TopoPodGroupManager, RoleBasedGroup, RoleSpec,
                             new_pod_group_manager)


class OraclePlacer:
    def __init__(self, topo):
        self.topo, self.n_nodes, self.blobs = topo, topo.n, []

    def score_assign(self, blob):
        self.blobs.append(np.array(blob))
        r = oracle_placer.place(self.topo, blob, want_matrix=False, want_topk=False)
        assert r["rc"] == 0
        return r["assign"], r["status"], r["domain"]


def mooncake(name="mc", gid=0, **kw):
    sh = synth.shape_mooncake()
    return RoleBasedGroup("default", name, [RoleSpec(r.name, r.replicas, tuple(r.deps), r.demand) for r in sh.roles],
                          gid=gid, policy_rules=sh.policy_rules, **kw)


def test_mooncake_levels_waves_and_feedback():
    topo = synth.make_topology(1024, seed=3, tiers=3)
    pl = OraclePlacer(topo)
    mgr = new_pod_group_manager("b200-topo", pl)
    out = mgr.reconcile_pod_groups_by_waves([mooncake()])[0]
    # dependencyOrder (pkg/dependency/dependency.go:129-205): master | decode, store, prefill | router
    assert len(pl.blobs) == 3
    assert [int(b[4]) for b in pl.blobs] == [1, 5, 1]          # replicas per wave
    assert [int(b[5]) for b in pl.blobs] == [1, 3, 1]          # role rows per wave
    assert list(out.nodes) == ["mc-mooncake-master-0", "mc-decode-0", "mc-mooncake-store-0", "mc-mooncake-store-1",
                               "mc-mooncake-store-2", "mc-prefill-0", "mc-router-0"]
    assert out.status == 0 and all(v >= 0 for v in out.nodes.values())
    # wave 2 carries every earlier placement as an anchor and the consumed capacity
    st = pl.blobs[2][8:24]
    anc = pl.blobs[2][st[8]: st[8] + 3 * st[7]].reshape(-1, 3)     # (node, role, count), aggregated
    assert anc[:, 2].sum() == 6 and st[9] >= 1
    # GetGroupSize = 7 (api/workloads/v1alpha2/helper.go:50-65)
    assert mgr.arith.group_size(mooncake().roles) == 7


def test_big_role_is_split_into_waves_of_32():
    topo = synth.make_topology(2048, seed=1, tiers=3)
    pl = OraclePlacer(topo)
    rbg = RoleBasedGroup("ns", "big", [RoleSpec("prefill", 70, (), 1), RoleSpec("decode", 3, (), 1)], gid=4)
    out = B200TopoPodGroupManager(pl).reconcile_pod_groups_by_waves([rbg])[0]
    # lexicographic inside the level: decode(3) then prefill(70): 3+29 | 32 | 9
    assert [int(b[4]) for b in pl.blobs] == [32, 32, 9]
    assert len(out.nodes) == 73 and list(out.nodes)[:4] == ["big-decode-0", "big-decode-1", "big-decode-2", "big-prefill-0"]


def test_coordination_targets_limit_the_pending_replicas():
    # rolebasedgroup_controller.go:509-518: role.Replicas is overridden by the scaling target
    topo = synth.make_topology(512, seed=2, tiers=2)
    pl = OraclePlacer(topo)
    rbg = RoleBasedGroup("ns", "pd", [RoleSpec("prefill", 300, (), 1), RoleSpec("decode", 100, (), 1)], gid=1,
                         targets={"prefill": 15, "decode": 5}, current={"prefill": 0, "decode": 0})
    out = B200TopoPodGroupManager(pl).reconcile_pod_groups_by_waves([rbg])[0]
    assert len(out.nodes) == 20 and "pd-prefill-14" in out.nodes and "pd-prefill-15" not in out.nodes
    # second round: current replicas shift the ordinals (stateful_instance_set_utils.go:74-76)
    rbg.targets, rbg.current = {"prefill": 30, "decode": 10}, {"prefill": 15, "decode": 5}
    out = B200TopoPodGroupManager(OraclePlacer(topo)).reconcile_pod_groups_by_waves([rbg])[0]
    assert sorted(out.nodes)[0] == "pd-decode-5" and "pd-prefill-29" in out.nodes and len(out.nodes) == 20


def test_gang_is_all_or_nothing_over_the_group():
    topo = synth.make_topology(256, seed=5, tiers=2, max_free=1)
    topo.free[:] = 0
    topo.free[7] = 1                      # room for exactly one pod in the whole cluster
    rbg = mooncake(annotations={GANG_SCHEDULING_KEY: "true"})
    out = B200TopoPodGroupManager(OraclePlacer(topo)).reconcile_pod_groups_by_waves([rbg])[0]
    assert out.status == 2 and len(out.nodes) == 7 and set(out.nodes.values()) == {-1}
    rbg2 = mooncake()                     # without gang the master (demand 0) and one pod still land
    out2 = B200TopoPodGroupManager(OraclePlacer(topo)).reconcile_pod_groups_by_waves([rbg2])[0]
    assert out2.status == 1 and sum(v >= 0 for v in out2.nodes.values()) >= 2


def test_exclusive_topology_keeps_the_group_in_one_domain():
    topo = synth.make_topology(1024, seed=8, tiers=3, owned_frac=0.3)
    ann = {EXCLUSIVE_TOPOLOGY_KEY: "kubernetes.io/hostname"}
    rbg = mooncake(gid=77, annotations=ann)
    rbg.roles[2].annotations = {ROLE_DISABLE_EXCLUSIVE_KEY: "true"}     # router opts out (annotation.go:29)
    pl = OraclePlacer(topo)
    mgr = B200TopoPodGroupManager(pl)
    out = mgr.reconcile_pod_groups_by_waves([rbg])[0]
    assert all(int(b[8 + 1]) & STEP_EXCLUSIVE for b in pl.blobs)
    doms = {int(topo.domain[n]) for k, n in out.nodes.items() if n >= 0 and "router" not in k}
    assert len(doms) == 1 and out.domain in doms
    assert topo.domain_owner[out.domain] in (-1, 77)
    assert int(pl.blobs[1][8 + 2]) == out.domain                     # later waves get the fixed domain
    tpl = {}
    mgr.InjectPodGroupLabels(rbg, tpl)
    hints = json.loads(tpl["metadata"]["annotations"][PLACEMENT_HINT_KEY])
    assert hints == {k: v for k, v in sorted(out.nodes.items()) if v >= 0}


def test_groups_blob_matches_the_wave_loop_inputs():
    """The GROUPS blob handed to rbgtopo_place_groups must describe the same
    problem the Python wave loop builds (same roles order, pair, anchors)."""
    topo = synth.make_topology(512, seed=4, tiers=3)
    rbgs = [mooncake(f"g{i}", gid=i, placed=[("prefill", 3 * i + 1)] * (i % 3)) for i in range(5)]
    mgr = B200TopoPodGroupManager(OraclePlacer(topo))
    blob, runs = mgr.groups_blob(rbgs)
    assert int(blob[0]) == 0x47474252 and int(blob[2]) == 5 and int(blob[4]) == 35
    rec = blob[8 + 12 * 2: 8 + 12 * 3]
    roles = blob[rec[4]: rec[4] + 4 * rec[3]].reshape(-1, 4)
    assert roles[:, 0].tolist() == [0, 1, 1, 1, 2] and roles[:, 1].tolist() == [1, 1, 3, 1, 1]
    anc = blob[rec[7]: rec[7] + 3 * rec[6]].reshape(-1, 3)         # 2 scheduled prefill pods on node 7, aggregated
    assert anc.tolist() == [[7, 3, 2]] and rec[8] == 14 and rec[9] == 7


def test_unknown_scheduler_name_is_rejected():
    # NewPodGroupManager, pkg/scheduler/podgroup_manager.go:82-92: unknown name -> error
    with pytest.raises(ValueError):
        new_pod_group_manager("kai", OraclePlacer(synth.make_topology(8, tiers=1)))


def test_scaling_rules_compute_the_targets_with_the_reference_arithmetic(golden):
    """rbg.scaling_rules + rbg.status -> targets exactly as CalculateScalingForAllCoordination
    (rolebasedgroup_controller.go:968-1054) would: every golden case of scaler_test.go through the
    plugin mirror, plus the minimum rule for a role paced by two rules."""
    from rbg_b200.plugin import HostArith, RoleStatus, ScalingRule
    arith = HostArith()
    for table in ("calculate_target_replicas", "progression_strategy"):
        for c in golden[table]["cases"]:
            if c.get("wantErr"):
                continue
            names = sorted(c["states"])
            rbg = RoleBasedGroup("ns", "g", [RoleSpec(nm, c["states"][nm][0]) for nm in names],
                                 scaling_rules=[ScalingRule(names, c["maxSkew"], c.get("progression", ""))],
                                 status={nm: RoleStatus(replicas=c["states"][nm][1], scheduled=c["states"][nm][2],
                                                        ready=c["states"][nm][3]) for nm in names})
            assert arith.scaling_targets(rbg) == c["want"], c["name"]
    rbg = RoleBasedGroup("ns", "g", [RoleSpec("a", 100), RoleSpec("b", 100), RoleSpec("c", 100)],
                         scaling_rules=[ScalingRule(["a", "b"], "10%"), ScalingRule(["b", "c"], "30%")])
    t = arith.scaling_targets(rbg)
    assert t["a"] == t["b"] and t["b"] < t["c"]         # b is paced by the tighter rule


def test_coordination_aware_batching_places_the_next_batch_ahead():
    """the batches the controller will walk through (MaxSkew-bounded steps),
    the current one placed and the next one pre-placed on top of it."""
    from rbg_b200.plugin import ScalingRule
    topo = synth.make_topology(512, seed=2, tiers=2)
    mgr = B200TopoPodGroupManager(OraclePlacer(topo))
    rbg = RoleBasedGroup("ns", "pd", [RoleSpec("prefill", 300, (), 1), RoleSpec("decode", 100, (), 1)], gid=1,
                         scaling_rules=[ScalingRule(["prefill", "decode"], "5%", "OrderScheduled")])
    batches = mgr.coordination_batches(rbg)
    assert batches[0] == {"prefill": 15, "decode": 5} and batches[1] == {"prefill": 30, "decode": 10}
    assert batches[-1] == {"prefill": 300, "decode": 100}
    assert all(b2[r] >= b1[r] for b1, b2 in zip(batches, batches[1:]) for r in b1)
    first, second = mgr.reconcile_ahead(rbg, 2, by_waves=True)
    assert len(first.nodes) == 20 and "pd-prefill-14" in first.nodes and "pd-prefill-15" not in first.nodes
    assert len(second.nodes) == 20 and sorted(second.nodes)[0] == "pd-decode-5" and "pd-prefill-29" in second.nodes
    # the pre-placed batch saw the first one: capacity taken by batch 1 is not handed out twice
    used = {}
    for p in (first, second):
        for node in p.nodes.values():
            if node >= 0:
                used[node] = used.get(node, 0) + 1
    assert all(cnt <= topo.free[node] for node, cnt in used.items())
    # one pass up to the second batch's targets, split by ordinal
    both = RoleBasedGroup("ns", "pd", rbg.roles, gid=1, targets=batches[1])
    want = B200TopoPodGroupManager(OraclePlacer(topo)).reconcile_pod_groups_by_waves([both])[0]
    assert {**first.nodes, **second.nodes} == want.nodes


def test_reconcile_ahead_with_a_role_no_rule_paces():
    """A role outside every ScalingRule is created whole by the current reconcile (its target is the
    spec's replica count): its replicas belong to batch 0, the paced roles are split by ordinal."""
    from rbg_b200.plugin import ScalingRule
    topo = synth.make_topology(512, seed=5, tiers=2)
    mgr = B200TopoPodGroupManager(OraclePlacer(topo))
    rbg = RoleBasedGroup("ns", "pd", [RoleSpec("prefill", 40, (), 1), RoleSpec("decode", 20, (), 1),
                                      RoleSpec("router", 2, (), 1)], gid=3,
                         scaling_rules=[ScalingRule(["prefill", "decode"], "25%", "OrderScheduled")])
    first, second = mgr.reconcile_ahead(rbg, 2, by_waves=True)
    assert "pd-router-0" in first.nodes and "pd-router-1" in first.nodes
    assert not any(k.startswith("pd-router-") for k in second.nodes)
    assert sum(k.startswith("pd-prefill-") for k in first.nodes) == 10
    assert sum(k.startswith("pd-prefill-") for k in second.nodes) == 10
    assert len(first.nodes) + len(second.nodes) == 2 + 20 + 10


def test_inject_pod_group_labels_keeps_the_wrapped_plugins_injection():
    """pkg/scheduler/podgroup_manager_test.go: the kube plugin labels the template, volcano annotates it,
    only when group-gang-scheduling == "true"; the placement hint rides beside it."""
    from rbg_b200.plugin import KUBE_POD_GROUP_LABEL, VOLCANO_GROUP_ANNOTATION
    topo = synth.make_topology(256, seed=1, tiers=2)
    gang = mooncake("mc", annotations={GANG_SCHEDULING_KEY: "true"})
    plain = mooncake("mc2", gid=1)
    for inner, where, key in (("scheduler-plugins", "labels", KUBE_POD_GROUP_LABEL), ("volcano", "annotations", VOLCANO_GROUP_ANNOTATION)):
        mgr = B200TopoPodGroupManager(OraclePlacer(topo), inner=inner)
        mgr.reconcile_pod_groups_by_waves([gang, plain])
        t1, t2 = {}, {}
        mgr.InjectPodGroupLabels(gang, t1)
        mgr.InjectPodGroupLabels(plain, t2)
        assert t1["metadata"][where][key] == "mc"
        assert key not in t2["metadata"].get(where, {})
        assert PLACEMENT_HINT_KEY in t1["metadata"]["annotations"] and PLACEMENT_HINT_KEY in t2["metadata"]["annotations"]
