"""Shared helpers of the GPU parity tests (test infrastructure)."""
import numpy as np

from oracle import placer as oracle_placer
from rbg_b200.engine import TopoPlacer


def new_engine(topo, **kw):
    eng = TopoPlacer(device=0, **kw)
    eng.set_topology(topo.row_ptr, topo.col_idx, topo.edge_w, topo.free, topo.domain, topo.domain_owner)
    return eng


def check_batch(eng, topo, blob, check_matrix=True, check_topk=True, ref=None):
    """Run blob on the GPU through the C ABI and on the oracle; assert bit-equality
    of the dense matrix, the final top-K keys, the assignment, status and domain."""
    if ref is None:
        ref = oracle_placer.place(topo, blob, want_matrix=check_matrix, want_topk=check_topk)
    assert ref["rc"] == 0, ref["rc"]
    h = eng.stage(blob)
    try:
        eng.run_staged(h, 1)
        assign, status, domain = eng.fetch(h)
        if check_matrix:
            for row in range(ref["matrix"].shape[0]):
                got = eng.read_scores(h, row)
                exp = ref["matrix"][row]
                if not np.array_equal(got.view(np.uint32), exp.view(np.uint32)):
                    bad = np.nonzero(got.view(np.uint32) != exp.view(np.uint32))[0]
                    raise AssertionError(f"matrix row {row}: {len(bad)} mismatches, first at node {bad[0]}: "
                                         f"gpu={got[bad[0]]} oracle={exp[bad[0]]}")
        if check_topk:
            for rr in range(ref["topk"].shape[0]):
                got = eng.read_topk(h, rr, 32)
                assert np.array_equal(got, ref["topk"][rr]), (rr, got[:8], ref["topk"][rr][:8])
        assert np.array_equal(assign, ref["assign"]), (assign[:32], ref["assign"][:32])
        assert np.array_equal(status, ref["status"])
        assert np.array_equal(domain, ref["domain"])
    finally:
        eng.release(h)
    # and through the host-buffer entry point
    a2, s2, d2 = eng.score_assign(blob)
    assert np.array_equal(a2, ref["assign"]) and np.array_equal(s2, ref["status"]) and np.array_equal(d2, ref["domain"])
    return ref


def plan_rows(gblob, topo):
    """(group, wave) -> dense-matrix row of the wave's first replica in a staged multi-wave plan
    (rows are in GROUP order: a group's waves are consecutive; rbgtopo_plan_describe column 4)."""
    from rbg_b200.engine import plan_steps
    return {(int(st[0]), int(st[1])): int(st[4]) for st in plan_steps(gblob, topo.n, len(topo.domain_owner))}
