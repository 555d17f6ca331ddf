"""CPU: Node objects -> snapshot arrays (rbg_b200/ingest.py): the CSR is
what rbgtopo_set_topology accepts (checked by the oracle's validator), independent of the order
the informer delivered the nodes, and placements run on it."""
import random

import numpy as np

from oracle import placer as oracle_placer
from rbg_b200.ingest import DEFAULT_TIER_LABELS, NodeInfo, build_topology, refresh
from rbg_b200.plugin import B200TopoPodGroupManager, RoleBasedGroup, RoleSpec
from test_plugin_host import OraclePlacer

NV, HOST, LEAF, ZONE = DEFAULT_TIER_LABELS


def cluster(n_domains=6, per_domain=8, gpus=8):
    nodes = []
    for d in range(n_domains):
        for k in range(per_domain):
            i = d * per_domain + k
            nodes.append(NodeInfo(f"node-{i:04d}", {NV: f"nvl-{d}", HOST: f"hg-{i // 4}", LEAF: f"leaf-{d // 2}",
                                                    ZONE: f"z-{d // 3}"},
                                  {"nvidia.com/gpu": gpus}, {"nvidia.com/gpu": i % 3}))
    return nodes


def test_csr_is_valid_symmetric_and_order_independent():
    nodes = cluster()
    topo, index = build_topology(nodes)
    assert oracle_placer.check_topology(topo) == 0
    shuffled = nodes[:]
    random.Random(3).shuffle(shuffled)
    topo2, index2 = build_topology(shuffled)
    assert index2.names == index.names == sorted(nd.name for nd in nodes)
    for f in ("row_ptr", "col_idx", "edge_w", "free", "domain", "domain_owner"):
        assert np.array_equal(getattr(topo, f), getattr(topo2, f)), f
    # closest shared tier decides the weight: node 0 and node 1 share the NVLink domain; 0 and 8 only the leaf
    row = lambda i: dict(zip(topo.col_idx[topo.row_ptr[i]:topo.row_ptr[i + 1]].tolist(),
                             topo.edge_w[topo.row_ptr[i]:topo.row_ptr[i + 1]].tolist()))
    assert row(0)[1] == 1000 and row(0)[7] == 1000
    assert row(0).get(8) == 10                      # other NVLink domain, same leaf (fan-out reaches it)
    assert all(w in (1000, 100, 10, 1) for w in topo.edge_w)
    assert topo.free[0] == 8 and topo.free[1] == 7 and topo.free[2] == 6
    assert len(index.domains) == 6 and topo.domain[0] == topo.domain[7] != topo.domain[8]


def test_unlabelled_and_cordoned_nodes():
    nodes = cluster(2, 4) + [NodeInfo("zz-plain", {}, {"nvidia.com/gpu": 4}),
                             NodeInfo("zz-cordoned", {NV: "nvl-0"}, {"nvidia.com/gpu": 8}, unschedulable=True)]
    topo, index = build_topology(nodes, domain_owner={"nvl-1": 7})
    assert oracle_placer.check_topology(topo) == 0
    plain = index.node_id("zz-plain")
    assert topo.row_ptr[plain + 1] == topo.row_ptr[plain]          # no label: no links, its own domain
    assert index.domains[topo.domain[plain]] == "node/zz-plain"
    assert topo.free[index.node_id("zz-cordoned")] == 0
    assert topo.domain_owner[index.domains.index("nvl-1")] == 7
    free, owner = refresh(topo, index, nodes, domain_owner={})
    assert np.array_equal(free, topo.free) and (owner == -1).all()


def test_placement_runs_on_an_ingested_snapshot():
    topo, index = build_topology(cluster(8, 8))
    rbg = RoleBasedGroup("ns", "pd", [RoleSpec("router", 1, (), 0), RoleSpec("prefill", 3, ("router",), 1),
                                      RoleSpec("decode", 4, ("router",), 1)], gid=1, policy_rules=[("prefill", "decode")])
    p = B200TopoPodGroupManager(OraclePlacer(topo)).reconcile_pod_groups_by_waves([rbg])[0]
    assert p.status == 0 and len(p.nodes) == 8
    doms = {topo.domain[v] for v in p.nodes.values()}
    assert len(doms) == 1                                          # the group packs into one NVLink domain
    assert all(index.names[v].startswith("node-") for v in p.nodes.values())
