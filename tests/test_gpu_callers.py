"""GPU: the callers either side of the hot path run on the CUDA engine and agree with the
same flow on the CPU oracle — an ingested Node snapshot (f1), coordination-aware batching (f4) and a
snapshot refresh from Node objects between reconciles."""
import numpy as np
import pytest

from rbg_b200.ingest import DEFAULT_TIER_LABELS, NodeInfo, build_topology, refresh
from rbg_b200.plugin import (EXCLUSIVE_TOPOLOGY_KEY, B200TopoPodGroupManager, RoleBasedGroup, RoleSpec,
                             ScalingRule)
from gpu_util import check_batch, new_engine
from test_plugin_host import OraclePlacer, mooncake

pytestmark = pytest.mark.gpu
NV, HOST, LEAF, ZONE = DEFAULT_TIER_LABELS


def cluster(n_domains, per_domain, gpus=8, used=lambda i: i % 3):
    return [NodeInfo(f"node-{d * per_domain + k:04d}",
                     {NV: f"nvl-{d}", HOST: f"hg-{(d * per_domain + k) // 4}", LEAF: f"leaf-{d // 2}", ZONE: f"z-{d // 3}"},
                     {"nvidia.com/gpu": gpus}, {"nvidia.com/gpu": used(d * per_domain + k)})
            for d in range(n_domains) for k in range(per_domain)]


def fleet(n):
    out = []
    for g in range(n):
        if g % 3 == 0:
            out.append(mooncake(f"mc{g}", gid=g))
        else:
            out.append(RoleBasedGroup("ns", f"pd{g}", [RoleSpec("router", 1, (), 0), RoleSpec("prefill", 2 + g % 3, ("router",), 1),
                                                       RoleSpec("decode", 4, ("router",), 1)], gid=g,
                                      policy_rules=[("prefill", "decode")],
                                      annotations={EXCLUSIVE_TOPOLOGY_KEY: "nvlink-domain"} if g % 4 == 1 else {}))
    return out


def test_ingested_snapshot_through_the_cuda_path():
    """Node objects -> build_topology -> rbgtopo_set_topology: step batches bit-equal to the oracle (matrix, lists,
    placements), a fleet placed through rbgtopo_place_groups equal to the oracle's wave loop, and again after the
    informer reported new allocations (refresh -> rbgtopo_update_nodes)."""
    nodes = cluster(24, 8)
    topo, index = build_topology(nodes, domain_owner={"nvl-3": 1})
    eng = new_engine(topo)
    try:
        rbgs = fleet(12)
        mgr_gpu, mgr_cpu = B200TopoPodGroupManager(eng), B200TopoPodGroupManager(OraclePlacer(topo))
        # one step batch of the wave loop, checked word for word (dense matrix rows included)
        pl = OraclePlacer(topo)
        B200TopoPodGroupManager(pl).reconcile_pod_groups_by_waves(rbgs[:4])
        for blob in pl.blobs[:3]:
            check_batch(eng, topo, blob)
        got = mgr_gpu.reconcile_pod_groups(rbgs)
        want = mgr_cpu.reconcile_pod_groups_by_waves(rbgs)
        assert [(p.status, p.nodes, p.domain) for p in got] == [(p.status, p.nodes, p.domain) for p in want]
        assert all(index.names[v].startswith("node-") for p in got for v in p.nodes.values() if v >= 0)
        # the informer reports pods bound elsewhere: capacities change, the CSR does not
        busier = cluster(24, 8, used=lambda i: (i * 7) % 9)
        free, owner = refresh(topo, index, busier, domain_owner={"nvl-3": 1})
        topo2 = type(topo)(topo.row_ptr, topo.col_idx, topo.edge_w, free, topo.domain, owner)
        eng.update_nodes(np.ascontiguousarray(free, dtype=np.int32))
        got2 = mgr_gpu.reconcile_pod_groups(rbgs)
        want2 = B200TopoPodGroupManager(OraclePlacer(topo2)).reconcile_pod_groups_by_waves(rbgs)
        assert [(p.status, p.nodes, p.domain) for p in got2] == [(p.status, p.nodes, p.domain) for p in want2]
        assert [p.nodes for p in got2] != [p.nodes for p in got]
    finally:
        eng.close()


def test_coordination_aware_batching_on_the_gpu():
    """reconcile_ahead (current batch placed, the next one pre-placed on top of it) gives the same hints on the CUDA
    engine as on the oracle placer, for a paced group and for one with an unpaced role."""
    from rbg_b200 import synth
    topo = synth.make_topology(512, seed=2, tiers=2)
    eng = new_engine(topo)
    try:
        for rbg in (
            RoleBasedGroup("ns", "pd", [RoleSpec("prefill", 300, (), 1), RoleSpec("decode", 100, (), 1)], gid=1,
                           scaling_rules=[ScalingRule(["prefill", "decode"], "5%", "OrderScheduled")]),
            RoleBasedGroup("ns", "pr", [RoleSpec("prefill", 40, (), 1), RoleSpec("decode", 20, (), 1), RoleSpec("router", 2, (), 1)],
                           gid=3, scaling_rules=[ScalingRule(["prefill", "decode"], "25%", "OrderScheduled")]),
        ):
            for by_waves in (True, False):
                a1, a2 = B200TopoPodGroupManager(eng).reconcile_ahead(rbg, 2, by_waves=by_waves)
                b1, b2 = B200TopoPodGroupManager(OraclePlacer(topo)).reconcile_ahead(rbg, 2, by_waves=True)
                assert (a1.nodes, a1.status) == (b1.nodes, b1.status), (rbg.name, by_waves)
                assert (a2.nodes, a2.status) == (b2.nodes, b2.status), (rbg.name, by_waves)
    finally:
        eng.close()
