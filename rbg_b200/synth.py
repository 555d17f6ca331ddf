"""Seeded synthetic inputs: tiered cluster topologies and
RoleBasedGroup fleets shaped like the reference's example manifests.

Everything derives from a counter-based SplitMix64 hash of (seed, stream, index)
so any language can regenerate identical inputs.  Pure numpy.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

TIER_W = (1000, 100, 10, 1)          # NVLink > PCIe > RDMA > VPC (README.md:53 order)
TIER_SIZE = (8, 32, 256, 1 << 30)    # NVLink domain, host group, RDMA leaf, VPC


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def rand_u64(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = splitmix64(np.asarray([seed * 0x1000193 + stream], dtype=np.uint64))[0]
        return splitmix64(np.asarray(idx, dtype=np.uint64) ^ base)


@dataclass
class Topology:
    row_ptr: np.ndarray
    col_idx: np.ndarray
    edge_w: np.ndarray
    free: np.ndarray
    domain: np.ndarray
    domain_owner: np.ndarray

    @property
    def n(self) -> int:
        return len(self.row_ptr) - 1

    @property
    def e(self) -> int:
        return len(self.col_idx)


def make_topology(n: int, seed: int = 0, tiers: int = 4, samples_per_tier: int = 4,
                  owned_frac: float = 0.0, max_free: int = 8) -> Topology:
    """N nodes in a `tiers`-level hierarchy.  Every node links to all peers of its
    NVLink domain (weight 1000) and to `samples_per_tier` sampled peers in each
    higher tier (100 / 10 / 1), made symmetric, rows sorted, no duplicates (a
    pair keeps its closest tier)."""
    with np.errstate(over="ignore"):
        ids = np.arange(n, dtype=np.int64)
        src: List[np.ndarray] = []
        dst: List[np.ndarray] = []
        # tier 0: NVLink clique
        g = TIER_SIZE[0]
        for off in range(1, g):
            d = (ids // g) * g + (ids % g + off) % g
            ok = d < n
            src.append(ids[ok]); dst.append(d[ok])
        for t in range(1, tiers):
            size = min(TIER_SIZE[t], n) if t < 3 else n
            inner = TIER_SIZE[t - 1]
            for s in range(samples_per_tier):
                r = rand_u64(seed, 10 * t + s, ids)
                start = (ids // size) * size
                span = np.minimum(start + size, n) - start
                d = start + (r % span.astype(np.uint64)).astype(np.int64)
                ok = (d // inner) != (ids // inner)       # must leave the lower tier
                src.append(ids[ok]); dst.append(d[ok])
        a = np.concatenate(src); b = np.concatenate(dst)
        lo = np.minimum(a, b); hi = np.maximum(a, b)
        pairs = np.unique(lo * n + hi)
        lo = pairs // n; hi = pairs % n
        u = np.concatenate([lo, hi]); v = np.concatenate([hi, lo])
        order = np.lexsort((v, u))
        u = u[order]; v = v[order]
        w = np.full(len(u), TIER_W[3], dtype=np.int32)
        for t in (2, 1, 0):
            same = (u // TIER_SIZE[t]) == (v // TIER_SIZE[t])
            w[same] = TIER_W[t]
        row_ptr = np.zeros(n + 1, dtype=np.int64)
        np.add.at(row_ptr, u + 1, 1)
        row_ptr = np.cumsum(row_ptr)
        free = (rand_u64(seed, 1, ids) % np.uint64(max_free + 1)).astype(np.int32)
        domain = (ids // TIER_SIZE[0]).astype(np.int32)
        nd = int(domain.max()) + 1
        owner = np.full(nd, -1, dtype=np.int32)
        if owned_frac > 0:
            r = rand_u64(seed, 2, np.arange(nd))
            taken = (r % np.uint64(1000)) < np.uint64(int(owned_frac * 1000))
            owner[taken] = (1_000_000 + np.arange(nd))[taken]   # foreign group ids
    return Topology(row_ptr.astype(np.int32), v.astype(np.int32), w, free, domain, owner)


# ---- RoleBasedGroup shapes taken from the reference's example manifests -------
@dataclass
class RoleShape:
    name: str
    replicas: int
    demand: int = 1
    deps: Sequence[str] = ()
    exclusive: bool = True    # False = role-disable-exclusive annotation


@dataclass
class GroupShape:
    name: str
    roles: List[RoleShape]
    exclusive: bool = False
    gang: bool = False
    policy_rules: List[Sequence[str]] = field(default_factory=list)  # CoordinatedPolicy role sets


def shape_sglang_pd() -> GroupShape:
    """examples/pd-disagg/sglang/sglang-pd.yaml: router/prefill/decode x (1,1,1)
    (:7,32,97; the router's dependency line :10 is commented out)."""
    return GroupShape("sglang-pd", [RoleShape("router", 1, 0), RoleShape("prefill", 1, 1),
                                    RoleShape("decode", 1, 1)],
                      policy_rules=[("prefill", "decode")])


def shape_pd_144() -> GroupShape:
    """BASELINE.json configs[1]: 3-role router/prefill/decode x (1,4,4), R = 9."""
    return GroupShape("pd-144", [RoleShape("router", 1, 0), RoleShape("prefill", 4, 1),
                                 RoleShape("decode", 4, 1)],
                      policy_rules=[("prefill", "decode")])


def shape_mooncake() -> GroupShape:
    """examples/mooncake/pd-disaggregated-with-mooncake.yaml: 5 roles, 7 pods
    (:7 master x1, :42 store x3 deps master, :83 router x1 deps prefill+decode,
    :104 prefill x1 deps master, :183 decode x1 deps master)."""
    return GroupShape("mooncake-pd", [
        RoleShape("mooncake-master", 1, 0),
        RoleShape("mooncake-store", 3, 1, ("mooncake-master",)),
        RoleShape("router", 1, 0, ("prefill", "decode")),
        RoleShape("prefill", 1, 1, ("mooncake-master",)),
        RoleShape("decode", 1, 1, ("mooncake-master",)),
    ], policy_rules=[("prefill", "decode")])


def shape_fleet8() -> GroupShape:
    """BASELINE.json configs[3]: router 1 / prefill 3 / decode 4 = 8 replicas."""
    return GroupShape("fleet8", [RoleShape("router", 1, 0), RoleShape("prefill", 3, 1),
                                 RoleShape("decode", 4, 1)],
                      policy_rules=[("prefill", "decode")])


def random_anchors(n_nodes: int, n_roles: int, count: int, seed: int, gid: int):
    """`count` already-placed pods (node, role, 1) for a partially deployed group."""
    idx = np.arange(count)
    nodes = (rand_u64(seed, 100 + gid, idx) % np.uint64(n_nodes)).astype(np.int64)
    roles = (rand_u64(seed, 200 + gid, idx) % np.uint64(max(n_roles, 1))).astype(np.int64)
    return [(int(a), int(b), 1) for a, b in zip(nodes, roles)]
