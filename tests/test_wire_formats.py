"""CPU: the two wire formats of include/rbgtopo.h as rbg_b200/blob.py writes them — header
words, record tables, section offsets (role sections 16-byte aligned), prefix fields — and that
the library's host-side validation accepts them (rbgtopo_plan_describe for GROUPS blobs)."""
import numpy as np

from rbg_b200.blob import (GROUP_WORDS, GROUPS_MAGIC, HDR_WORDS, MAGIC, STEP_WORDS, BlobBuilder, Group, GroupsBuilder,
                           Step, blob_totals, tile_groups_blob)
from test_plan_describe import describe


def test_step_blob_layout():
    bb = BlobBuilder()
    bb.add(Step(gid=3, roles=[(2, 1, 5, 1), (1, 0, 2, 0)], pair=[[1, 0, 1], [0, 1, 1]], anchors=[(7, 2, 1)],
                consumed=[(7, 2), (9, 1)], flags=3, fixed_domain=4))
    bb.add(Step(gid=5, roles=[(4, 2, 16, 1)], pair=[[1]]))
    b = bb.build()
    assert b.dtype == np.int32 and b[0] == MAGIC and b[1] == 1 and b[2] == 2 and b[3] == len(b)
    assert blob_totals(b) == (2, 7, 3) and b[6] == 0 and b[7] == 0
    s0 = b[HDR_WORDS:HDR_WORDS + STEP_WORDS]
    s1 = b[HDR_WORDS + STEP_WORDS:HDR_WORDS + 2 * STEP_WORDS]
    assert list(s0[:4]) == [3, 3, 4, 2] and s0[5] == 3 and s0[7] == 1 and s0[9] == 2 and s0[11] == 3
    assert list(s0[12:16]) == [0, 0, 0, 0] and list(s1[12:16]) == [3, 2, 0, 0]
    for s in (s0, s1):
        assert s[4] % 4 == 0 and s[4] >= HDR_WORDS + 2 * STEP_WORDS          # role records: 16-byte aligned
        assert s[6] == s[4] + 4 * s[3] and s[8] == s[6] + s[3] * s[5] and s[10] == s[8] + 3 * s[7]
    assert list(b[s0[4]:s0[4] + 8]) == [2, 1, 5, 1, 1, 0, 2, 0]
    assert list(b[s0[6]:s0[6] + 6]) == [1, 0, 1, 0, 1, 1]
    assert list(b[s0[8]:s0[8] + 3]) == [7, 2, 1] and list(b[s0[10]:s0[10] + 4]) == [7, 2, 9, 1]
    assert s1[10] + 2 * s1[9] == len(b)
    assert len(BlobBuilder().build()) == HDR_WORDS and blob_totals(BlobBuilder().build()) == (0, 0, 0)


def test_groups_blob_layout_and_tiling():
    g = Group(gid=10, roles=[(0, 1, 0, 1), (1, 3, 1, 1), (1, 2, 1, 0)], pair=[[1, 1, 1], [1, 1, 0], [1, 0, 1]],
              anchors=[(5, 0, 1)], flags=2, fixed_domain=-1)
    b = GroupsBuilder().add(g).add(Group(gid=11, roles=[(0, 4, 1, 1)], pair=[[1]])).build()
    assert b[0] == GROUPS_MAGIC and b[2] == 2 and b[3] == len(b) and b[4] == 10
    r0 = b[HDR_WORDS:HDR_WORDS + GROUP_WORDS]
    r1 = b[HDR_WORDS + GROUP_WORDS:HDR_WORDS + 2 * GROUP_WORDS]
    assert list(r0[:4]) == [10, 2, -1, 3] and r0[6] == 1 and list(r0[8:10]) == [0, 6]
    assert r0[5] == r0[4] + 12 and r0[7] == r0[5] + 9 and r1[4] == r0[7] + 3
    assert list(r1[:4]) == [11, 0, -1, 1] and list(r1[8:10]) == [6, 4]
    rc, steps, n_waves, _ = describe(b, n_nodes=64, n_domains=4)
    assert rc == 0 and n_waves == 2 and [(int(s[0]), int(s[1])) for s in steps] == [(0, 0), (1, 0), (0, 1)]
    # tiling one group into a fleet: gids step, offsets shift, assign offsets are the prefix of pending
    one = GroupsBuilder().add(g).build()
    fleet = tile_groups_blob(one, 5, gid_stride=2)
    assert fleet[2] == 5 and fleet[3] == len(fleet) and fleet[4] == 30
    recs = fleet[HDR_WORDS:HDR_WORDS + 5 * GROUP_WORDS].reshape(5, GROUP_WORDS)
    assert list(recs[:, 0]) == [10, 12, 14, 16, 18] and list(recs[:, 8]) == [0, 6, 12, 18, 24]
    same = GroupsBuilder()
    for i in range(5):
        same.add(Group(gid=10 + 2 * i, roles=g.roles, pair=g.pair, anchors=g.anchors, flags=g.flags))
    assert np.array_equal(fleet, same.build())
    rc, steps, n_waves, _ = describe(fleet, n_nodes=64, n_domains=4)
    assert rc == 0 and n_waves == 2 and len(steps) == 10
