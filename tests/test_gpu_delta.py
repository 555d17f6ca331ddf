"""GPU: incremental snapshot refresh (rbgtopo_update_nodes_delta).  After every
delta the device-resident base vector and background order must be exactly what a full refresh of the
same capacities gives: checked through the dense matrix bits, the top-K key lists and the placements of
step batches and whole fleets against the CPU oracle run on the updated snapshot."""
import numpy as np
import pytest

from rbg_b200 import synth
from rbg_b200.plugin import B200TopoPodGroupManager

pytestmark = pytest.mark.gpu


def _check(eng, topo, seed, fleet):
    from gpu_util import check_batch
    from test_gpu_groups import _oracle_manager
    from test_gpu_parity import _random_steps
    check_batch(eng, topo, _random_steps(topo, seed, 16, excl=False, gang=True))
    got = B200TopoPodGroupManager(eng).reconcile_pod_groups(fleet)
    ref = _oracle_manager(topo).reconcile_pod_groups_by_waves(fleet)
    for a, c in zip(got, ref):
        assert a.nodes == c.nodes and a.status == c.status and a.domain == c.domain


def test_small_deltas_match_a_full_refresh():
    from gpu_util import new_engine
    from test_gpu_groups import _fleet
    n = 6000
    topo = synth.make_topology(n, seed=31, tiers=4)
    fleet = _fleet(n, 12, seed=4)
    eng = new_engine(topo)
    rng = np.random.default_rng(9)
    for step in range(10):
        k = int(rng.integers(1, 24))
        nodes = rng.choice(n, size=k, replace=False).astype(np.int32)
        vals = rng.integers(0, 13, size=k).astype(np.int32)      # above F = 8 too: the scores only see min(free, 8)
        if step % 3 == 0:                                        # duplicates: the last value wins
            nodes = np.concatenate([nodes, nodes[:2]])
            vals = np.concatenate([vals, np.array([1, 7], dtype=np.int32)[:len(nodes[:2])]])
        eng.update_nodes_delta(nodes, vals, generation=100 + step)
        for nd, v in zip(nodes, vals):
            topo.free[nd] = v
        _check(eng, topo, 200 + step, fleet)
    assert eng.stats()["generation"] == 109
    eng.close()


def test_delta_then_full_then_delta_and_the_large_delta_fallback():
    from gpu_util import new_engine
    from test_gpu_groups import _fleet
    n = 3000
    topo = synth.make_topology(n, seed=5, tiers=3)
    fleet = _fleet(n, 8, seed=6)
    eng = new_engine(topo)
    rng = np.random.default_rng(1)
    # small delta
    nodes = rng.choice(n, size=5, replace=False).astype(np.int32)
    eng.update_nodes_delta(nodes, np.zeros(5, dtype=np.int32))
    topo.free[nodes] = 0
    _check(eng, topo, 1, fleet)
    # full refresh in between
    topo.free = rng.integers(0, 9, size=n).astype(np.int32)
    eng.update_nodes(topo.free)
    _check(eng, topo, 2, fleet)
    # 10 % of the nodes at once: more neighbourhoods than one repair handles -> the library refreshes fully
    nodes = rng.choice(n, size=n // 10, replace=False).astype(np.int32)
    vals = rng.integers(0, 9, size=len(nodes)).astype(np.int32)
    eng.update_nodes_delta(nodes, vals)
    topo.free[nodes] = vals
    _check(eng, topo, 3, fleet)
    # and small again on top of it
    eng.update_nodes_delta(np.array([7, 8, 9], dtype=np.int32), np.array([8, 0, 3], dtype=np.int32))
    topo.free[[7, 8, 9]] = [8, 0, 3]
    _check(eng, topo, 4, fleet)
    # bad input: codes, not crashes
    from rbg_b200.engine import RbgTopoError
    with pytest.raises(RbgTopoError):
        eng.update_nodes_delta(np.array([n], dtype=np.int32), np.array([1], dtype=np.int32))
    with pytest.raises(RbgTopoError):
        eng.update_nodes_delta(np.array([0], dtype=np.int32), np.array([-1], dtype=np.int32))
    eng.close()
