"""ctypes access to the C oracle (oracle/placer_oracle.c).  TEST INFRASTRUCTURE:
imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Parity is UNPINNED upstream (see the C file's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
KMAX = 32
_lib = None


def build() -> str:
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_place.restype = C.c_int
        _lib.oracle_place_fast.restype = C.c_int
        _lib.oracle_check_topology.restype = C.c_int
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _p(a, t=C.c_int32):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def max_threads() -> int:
    return load().oracle_max_threads()


def check_topology(topo) -> int:
    rp, ci, ew = _i32(topo.row_ptr), _i32(topo.col_idx), _i32(topo.edge_w)
    fr, dm, ow = _i32(topo.free), _i32(topo.domain), _i32(topo.domain_owner)
    return load().oracle_check_topology(C.c_int32(len(rp) - 1), C.c_int64(len(ci)), _p(rp), _p(ci), _p(ew),
                                        _p(fr), _p(dm), C.c_int32(len(ow)), _p(ow))


_MATRIX_CACHE = {}


def place_fast(topo, blob, want_matrix=True, want_topk=True, nthreads=1, reuse_matrix=False):
    """The CPU variant with the GPU path's algebra (oracle/placer_fast.c): same contract and, by
    the exactness contract of the spec, the same bits as place()."""
    return place(topo, blob, want_matrix, want_topk, nthreads, reuse_matrix, _fn="oracle_place_fast")


def place(topo, blob, want_matrix=True, want_topk=True, nthreads=1, reuse_matrix=False, _fn="oracle_place"):
    """Run the oracle on a batch.  Returns dict(rc, assign, status, domain,
    matrix [total R][N] or None, topk [rolerows][32] or None).
    reuse_matrix: write the dense matrix into a buffer kept per shape instead of a fresh array
    (timing runs: a fresh 50 MB array per call is page-fault time, not oracle time); the returned
    matrix is then only valid until the next call with the same shape."""
    lib = load()
    rp, ci, ew = _i32(topo.row_ptr), _i32(topo.col_idx), _i32(topo.edge_w)
    fr, dm, ow = _i32(topo.free), _i32(topo.domain), _i32(topo.domain_owner)
    blob = _i32(blob)
    n = len(rp) - 1
    ns, tr, tp = int(blob[2]), int(blob[4]), int(blob[5])
    matrix = None
    if want_matrix:
        shape = (max(tr, 1), n)
        if reuse_matrix:
            matrix = _MATRIX_CACHE.get(shape)
            if matrix is None:
                matrix = _MATRIX_CACHE[shape] = np.empty(shape, dtype=np.float32)
        else:
            matrix = np.empty(shape, dtype=np.float32)
    topk = np.zeros((max(tp, 1), KMAX), dtype=np.uint64) if want_topk else None
    assign = np.full(max(tr, 1), -2, dtype=np.int32)
    status = np.full(max(ns, 1), -2, dtype=np.int32)
    domain = np.full(max(ns, 1), -2, dtype=np.int32)
    rc = getattr(lib, _fn)(C.c_int32(n), C.c_int64(len(ci)), _p(rp), _p(ci), _p(ew), _p(fr), _p(dm),
                          C.c_int32(len(ow)), _p(ow), _p(blob), C.c_int64(len(blob)),
                          _p(matrix, C.c_float), _p(topk, C.c_uint64), _p(assign), _p(status), _p(domain),
                          C.c_int32(nthreads))
    return dict(rc=rc, assign=assign[:tr], status=status[:ns], domain=domain[:ns],
                matrix=None if matrix is None else matrix[:tr], topk=None if topk is None else topk[:tp])
