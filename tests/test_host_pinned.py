"""The product's host-side mirrors of the reference-pinned arithmetic (C, exported
through the ABI: rbgtopo_* in include/rbgtopo.h) against the reference's golden
tables and, on sweeps, against oracle/refpinned.py."""
import ctypes as C
import math

import pytest

from oracle import refpinned as rp
from rbg_b200 import _lib
from rbg_b200.plugin import HostArith, RoleSpec


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def _arr(v):
    return (C.c_int32 * len(v))(*v)


PROG = {"": 0, None: 0, "OrderScheduled": 1, "OrderReady": 2}


def _ctr(lib, max_skew, prog, states, order):
    n = len(order)
    tgt = (C.c_int32 * n)()
    rc = lib.rbgtopo_calculate_target_replicas(max_skew, PROG[prog], n, _arr([states[r][0] for r in order]),
                                               _arr([states[r][1] for r in order]), _arr([states[r][2] for r in order]),
                                               _arr([states[r][3] for r in order]), tgt)
    return rc, dict(zip(order, tgt))


def test_calculate_target_replicas_golden(lib, golden):
    # pkg/coordination/coordinationscaling/scaler_test.go:100-518, :599-759
    for table in ("calculate_target_replicas", "progression_strategy"):
        for c in golden[table]["cases"]:
            out = C.c_double()
            assert lib.rbgtopo_parse_percentage(c["maxSkew"].encode(), C.byref(out)) == 0
            if c.get("wantErr"):
                assert _ctr(lib, out.value, "", {}, [])[0] != 0
                continue
            for order in (sorted(c["states"]), sorted(c["states"], reverse=True)):
                rc, got = _ctr(lib, out.value, c.get("progression", ""), c["states"], order)
                assert rc == 0 and got == c["want"], c["name"]


def test_parse_percentage_golden(lib, golden):
    # scaler_test.go:520-597
    for c in golden["parse_percentage"]["cases"]:
        out = C.c_double()
        rc = lib.rbgtopo_parse_percentage(c["in"].encode(), C.byref(out))
        if c.get("wantErr"):
            assert rc != 0, c["in"]
        else:
            assert rc == 0 and out.value == c["want"], c["in"]


def test_updated_replicas_bound_golden(lib, golden):
    # rolebasedgroup_controller_test.go:1283-1377
    for c in golden["updated_replicas_bound"]["cases"]:
        lo, hi = C.c_int32(), C.c_int32()
        s = int(c["maxSkew"].rstrip("%"))
        assert lib.rbgtopo_updated_replicas_bound(s, c["refUpdated"], c["refDesired"], c["requestDesired"],
                                                  C.byref(lo), C.byref(hi)) == 0
        assert (lo.value, hi.value) == (c["lower"], c["upper"]), c["name"]


def test_scaled_value_matches_intstr(lib):
    # vendor/k8s.io/apimachinery/pkg/util/intstr/intstr.go:181-197
    for pct in (0, 1, 5, 10, 33, 50, 99, 100, 150):
        for total in (0, 1, 2, 3, 7, 10, 99, 100, 1000):
            for up in (0, 1):
                assert lib.rbgtopo_scaled_value(1, pct, total, up) == \
                    rp.get_scaled_value_from_int_or_percent(f"{pct}%", total, bool(up))
    assert lib.rbgtopo_scaled_value(0, 7, 100, 1) == 7


def _roll(lib, s_pct, desired, updated):
    """Test_CalculateNextRollingTarget loop (rolebasedgroup_controller_test.go:207-250) via the C mirror,
    cross-checked step by step against oracle/refpinned.py."""
    names = sorted(desired)
    d = [desired[k] for k in names]
    u = [updated[k] for k in names]
    bias = max(int(math.ceil(10000.0 / x)) for x in d)
    max_bp = max(1, int(math.ceil(s_pct * 10000 / 100)))
    for _ in range(100000):
        tgt = (C.c_int32 * len(d))()
        rc = lib.rbgtopo_next_rolling_target(s_pct, len(d), _arr(d), _arr(u), _arr(d), tgt)
        ref = rp.calculate_next_rolling_target(f"{s_pct}%", names, dict(zip(names, d)), dict(zip(names, u)),
                                               dict(zip(names, d)))
        assert rc == 0 and list(tgt) == [ref[k] for k in names], (d, u, list(tgt), ref)
        u = list(tgt)
        ratios = [u[i] / d[i] for i in range(len(d))]
        assert int(math.ceil(10000.0 * (max(ratios) - min(ratios)))) <= bias + max_bp, (d, u)
        if all(u[i] >= d[i] for i in range(len(d))):
            return
    raise AssertionError("did not terminate")


def test_next_rolling_target(lib, golden):
    # rolebasedgroup_controller_test.go:54-332 (seeds + thinned sweeps; full sweeps in test_refpinned_golden)
    for c in golden["next_rolling_target_seeds"]["cases"]:
        _roll(lib, int(c["maxSkew"].rstrip("%")), c["desired"], c["updated"])
    for p in range(1, 100, 7):
        for d in range(1, 100, 5):
            _roll(lib, 1, {"prefill": p, "decode": d}, {"prefill": 0, "decode": 0})


def test_dependency_levels_and_group_size(golden):
    # pkg/dependency/dependency_test.go:37-121; api/workloads/v1alpha2/helper.go:50-65
    ha = HostArith()
    for c in golden["dependency_order"]["cases"]:
        roles = [RoleSpec(n, 1, tuple(deps)) for n, deps in c["deps"].items()]
        if c.get("wantErr"):
            with pytest.raises(ValueError):
                ha.dependency_levels(roles)
        else:
            got = [[roles[i].name for i in lvl] for lvl in ha.dependency_levels(roles)]
            assert got == c["want"], c["name"]
    for c in golden["group_size"]["cases"]:
        roles = [RoleSpec(f"r{i}", r["replicas"], lws_size=(r.get("lws_size") or 0) if r.get("lws") else 0)
                 for i, r in enumerate(c["roles"])]
        assert ha.group_size(roles) == c["want"]


def test_dependency_levels_random_vs_oracle():
    import random
    ha = HostArith()
    rnd = random.Random(5)
    for _ in range(200):
        n = rnd.randint(1, 9)
        names = [f"role{rnd.randint(0, 99):02d}x{i}" for i in range(n)]
        deps = {nm: [names[j] for j in range(n) if j < i and rnd.random() < 0.3] for i, nm in enumerate(names)}
        roles = [RoleSpec(nm, 1, tuple(deps[nm])) for nm in names]
        rnd.shuffle(roles)
        got = [[roles[i].name for i in lvl] for lvl in ha.dependency_levels(roles)]
        assert got == rp.dependency_order(deps)


def _intstr(v):
    """IntOrString of a golden case -> (has, is_percent, value), or None when the string is one the
    Go parser rejects (those cases test intstr itself, not the functions above it)."""
    if v is None:
        return (0, 0, 0)
    if isinstance(v, int):
        return (1, 0, v)
    s = str(v)
    if s.endswith("%") and s[:-1].lstrip("-").isdigit():
        return (1, 1, int(s[:-1]))
    return None


def test_partition_replicas_and_non_zero(lib, golden):
    # pkg/utils/utils_test.go:342-482 and :484-583, through the C ABI
    out = C.c_int32()
    n = 0
    for c in golden["calculate_partition_replicas"]["cases"]:
        p = _intstr(c["partition"])
        if c.get("wantErr") or p is None:
            continue
        reps = -1 if c["replicas"] is None else c["replicas"]
        assert lib.rbgtopo_partition_replicas(p[0], p[1], p[2], reps, C.byref(out)) == 0
        assert out.value == c["want"], c["name"]
        assert out.value == rp.calculate_partition_replicas(c["partition"], c["replicas"])
        n += 1
    assert n >= 8
    n = 0
    for c in golden["parse_intstr_as_non_zero"]["cases"]:
        p = _intstr(c["in"])
        if c.get("wantErr") or p is None:
            continue
        assert lib.rbgtopo_intstr_non_zero(p[1], p[2], c["replicas"], C.byref(out)) == 0
        assert out.value == c["want"], c["name"]
        n += 1
    assert n >= 5


def test_merge_rolling_update(lib, golden):
    # rolebasedgroup_controller_test.go:1090-1281 (Test_mergeStrategyRollingUpdate), role by role through the C ABI
    def pack(st):
        mu, pt = _intstr(st[0]), _intstr(st[1])
        return None if mu is None or pt is None else list(mu) + list(pt)

    def unpack(v):
        def one(h, pct, val):
            return None if not h else (f"{val}%" if pct else val)
        return [one(*v[0:3]), one(*v[3:6])]
    n = 0
    for c in golden["merge_strategy_rolling_update"]["cases"]:
        for role, want in c["want"].items():
            if role not in c["a"] or role not in c["b"]:
                assert want == (c["a"].get(role) or c["b"].get(role))     # present in one map only: taken as it is
                continue
            a, b = pack(c["a"][role]), pack(c["b"][role])
            if a is None or b is None:
                continue
            out = (C.c_int32 * 6)()
            assert lib.rbgtopo_merge_rolling_update((C.c_int32 * 6)(*a), (C.c_int32 * 6)(*b), out) == 0
            assert unpack(list(out)) == want, c["name"]
            n += 1
    assert n >= 3


def test_naming_keys_and_annotation_inheritance(lib, golden):
    # helper.go:68-81 (63-byte cut + TrimRight "-"), :135-144 (sha1), annotation_inheritance_test.go:25-49 — product C functions
    import hashlib
    buf = C.create_string_buffer(64)
    for rbg, role in [("rbg", "prefill"), ("a" * 62, "-x"), ("x" * 40, "y" * 40), ("n", "r")]:
        n = lib.rbgtopo_workload_name(rbg.encode(), role.encode(), buf, 64)
        assert buf.value.decode() == rp.get_workload_name(rbg, role) and n == len(buf.value)
    key = C.create_string_buffer(41)
    for ns, name in [("default", "test-rbg"), ("ns", "n"), ("a" * 70, "b" * 70), ("", "")]:
        assert lib.rbgtopo_group_unique_key(ns.encode(), name.encode(), key, 41) == 40
        assert key.value.decode() == hashlib.sha1(f"{ns}/{name}".encode()).hexdigest() == rp.gen_group_unique_key(ns, name)
    for c in golden["inherit_pod_group_annotations"]["cases"]:
        ann, prefixes = c["annotations"] or {}, c["prefixes"]
        arr = (C.c_char_p * max(len(prefixes), 1))(*[p.encode() for p in prefixes])
        got = {k: v for k, v in ann.items() if lib.rbgtopo_inherits_annotation(k.encode(), len(prefixes), arr) == 1}
        assert (got or None) == c["want"]
