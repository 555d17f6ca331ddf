"""Builder of the batch wire format of include/rbgtopo.h (BLOB section)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

MAGIC = 0x54474252
VERSION = 1
HDR_WORDS = 8
STEP_WORDS = 16
STEP_EXCLUSIVE = 1
STEP_GANG = 2
ROLE_EXCLUSIVE = 1
MAX_STEP_ROLES = 8
MAX_STEP_REPLICAS = 32
MAX_GROUP_ROLES = 16
NEED_CAP = 16


@dataclass
class Step:
    """One wave of one dependency level of one RoleBasedGroup."""
    gid: int
    roles: List[Tuple[int, int, int, int]]          # (count, demand, need, role_flags)
    pair: Sequence[Sequence[int]] = field(default_factory=list)  # [P][Q]
    anchors: List[Tuple[int, int, int]] = field(default_factory=list)   # (node, role q, count)
    consumed: List[Tuple[int, int]] = field(default_factory=list)       # (node, amount)
    flags: int = 0
    fixed_domain: int = -1

    @property
    def n_replicas(self) -> int:
        return sum(r[0] for r in self.roles)


class BlobBuilder:
    def __init__(self) -> None:
        self.steps: List[Step] = []

    def add(self, step: Step) -> "BlobBuilder":
        self.steps.append(step)
        return self

    def build(self) -> np.ndarray:
        ns = len(self.steps)
        body: List[int] = []
        table = np.zeros((ns, STEP_WORDS), dtype=np.int64)
        base = HDR_WORDS + ns * STEP_WORDS
        racc = pacc = 0
        for i, s in enumerate(self.steps):
            P = len(s.roles)
            Q = len(s.pair[0]) if (len(s.pair) and len(s.pair[0])) else 0
            role_off = base + len(body)
            for r in s.roles:
                body.extend(int(x) for x in r)
            pair_off = base + len(body)
            for p in range(P):
                row = s.pair[p] if Q else []
                assert len(row) == Q
                body.extend(int(x) for x in row)
            anchor_off = base + len(body)
            for a in s.anchors:
                body.extend(int(x) for x in a)
            cons_off = base + len(body)
            for c in s.consumed:
                body.extend(int(x) for x in c)
            R = s.n_replicas
            table[i] = [s.gid, s.flags, s.fixed_domain, P, role_off, Q, pair_off, len(s.anchors),
                        anchor_off, len(s.consumed), cons_off, R, racc, pacc, 0, 0]
            racc += R
            pacc += P
        words = base + len(body)
        out = np.zeros(words, dtype=np.int32)
        out[0:8] = [MAGIC, VERSION, ns, words, racc, pacc, 0, 0]
        out[HDR_WORDS:base] = table.reshape(-1)
        out[base:] = np.asarray(body, dtype=np.int64) if body else []
        return out


def blob_totals(blob: np.ndarray) -> Tuple[int, int, int]:
    """(n_steps, total replicas, total role rows) of a built blob."""
    return int(blob[2]), int(blob[4]), int(blob[5])
