"""Second, independent restatement of the frozen placement spec in
numpy / plain Python — TEST INFRASTRUCTURE.  It exists only to cross-check the C
oracle (oracle/placer_oracle.c) on small cases: two restatements written from the
spec text that agree are the best pin available for a path the reference does not
have (parity unpinned upstream).  Integer arithmetic throughout
(the spec's exactness contract makes the fp32 sums exact integers)."""
from __future__ import annotations

import numpy as np

F_CAP, SELF_W, KMAX = 8, 8000, 32
STEP_EXCLUSIVE, STEP_GANG, ROLE_EXCLUSIVE = 1, 2, 1


def _key(score: int, node: int) -> int:
    # orderable_u32 of a non-negative fp32 integer value = bits | 0x80000000
    bits = int(np.float32(score).view(np.uint32))
    return ((bits ^ 0x80000000) << 32) | (0xFFFFFFFF - node)


def place(topo, blob):
    """Returns dict(matrix [R][N] float32, topk [rolerows][32] uint64, assign, status, domain)."""
    blob = np.asarray(blob, dtype=np.int64)
    n = len(topo.row_ptr) - 1
    rp, ci, ew = (np.asarray(x, dtype=np.int64) for x in (topo.row_ptr, topo.col_idx, topo.edge_w))
    free = np.asarray(topo.free, dtype=np.int64)
    dom = np.asarray(topo.domain, dtype=np.int64)
    owner = np.asarray(topo.domain_owner, dtype=np.int64)
    rows_of = np.repeat(np.arange(n), np.diff(rp))
    ns, tr, tp = int(blob[2]), int(blob[4]), int(blob[5])
    matrix = np.full((tr, n), -np.inf, dtype=np.float32)
    topk = np.zeros((tp, KMAX), dtype=np.uint64)
    assign = np.full(tr, -1, dtype=np.int32)
    status = np.zeros(ns, dtype=np.int32)
    domain_out = np.full(ns, -1, dtype=np.int32)
    for s in range(ns):
        st = blob[8 + 16 * s: 8 + 16 * (s + 1)]
        gid, flags, fixed, P, role_off, Q, pair_off, na, anc_off, nc, cons_off, R, rep_off, rr_off = (int(x) for x in st[:14])
        roles = blob[role_off: role_off + 4 * P].reshape(P, 4)
        pair = blob[pair_off: pair_off + P * Q].reshape(P, Q) if Q else np.zeros((P, 0), dtype=np.int64)
        anchor = np.zeros((max(Q, 1), n), dtype=np.int64)
        for a in blob[anc_off: anc_off + 3 * na].reshape(na, 3):
            anchor[a[1], a[0]] += a[2]
        cons = np.zeros(n, dtype=np.int64)
        for c in blob[cons_off: cons_off + 2 * nc].reshape(nc, 2):
            cons[c[0]] += c[1]
        excl_step = bool(flags & STEP_EXCLUSIVE)
        S = []
        for p in range(P):
            cnt, demand, need, rflags = (int(x) for x in roles[p])
            A = need * np.minimum(free, F_CAP)
            for q in range(Q):
                A = A + pair[p, q] * anchor[q]
            sc = np.zeros(n, dtype=np.int64)
            np.add.at(sc, rows_of, ew * A[ci])
            sc += SELF_W * A
            assert sc.max(initial=0) < 1 << 24, "exactness contract"
            feas = (free - cons) >= demand
            if excl_step and (rflags & ROLE_EXCLUSIVE):
                o = owner[dom]
                feas &= (o == -1) | (o == gid)
            S.append((sc, feas))
        r = rep_off
        for p in range(P):
            row = np.where(S[p][1], S[p][0].astype(np.float32), np.float32(-np.inf))
            for _ in range(int(roles[p, 0])):
                matrix[r] = row
                r += 1
        dstar = -1
        if excl_step:
            if fixed >= 0:
                dstar = fixed
            else:
                for p in range(P):
                    if roles[p, 3] & ROLE_EXCLUSIVE:
                        sc, feas = S[p]
                        idx = np.nonzero(feas)[0]
                        if len(idx):
                            best = max(idx, key=lambda i: _key(int(sc[i]), int(i)))
                            dstar = int(dom[best])
                        break
        domain_out[s] = dstar
        lists, kacc = [], 0
        for p in range(P):
            kacc += int(roles[p, 0])
            K = min(kacc, n)
            sc, feas = S[p]
            ok = feas.copy()
            if excl_step and (roles[p, 3] & ROLE_EXCLUSIVE):
                ok &= dom == dstar
            keys = sorted((_key(int(sc[i]), int(i)) for i in np.nonzero(ok)[0]), reverse=True)[:K]
            lists.append(keys)
            topk[rr_off + p, :len(keys)] = np.array(keys, dtype=np.uint64)
        avail = free - cons
        r, unplaced = rep_off, 0
        for p in range(P):
            cnt, demand = int(roles[p, 0]), int(roles[p, 1])
            for _ in range(cnt):
                pick = -1
                for k in lists[p]:
                    node = 0xFFFFFFFF - (k & 0xFFFFFFFF)
                    if avail[node] >= demand:
                        pick = node
                        break
                if pick >= 0:
                    avail[pick] -= demand
                else:
                    unplaced += 1
                assign[r] = pick
                r += 1
        if unplaced and (flags & STEP_GANG):
            assign[rep_off: rep_off + R] = -1
            status[s] = 2
        else:
            status[s] = 1 if unplaced else 0
    return dict(matrix=matrix, topk=topk, assign=assign, status=status, domain=domain_out)
