"""Host-only timing of the plan geometry (rbgtopo_plan_describe = the host part of
rbgtopo_place_groups) for the bench fleet: no GPU needed.  RBGTOPO_HOST_THREADS / RBGTOPO_PROFILE_HOST apply."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from rbg_b200 import _lib, synth  # noqa: E402
from rbg_b200.plugin import B200TopoPodGroupManager  # noqa: E402
from test_plugin_host import OraclePlacer  # noqa: E402

n = 10000
topo = synth.make_topology(n, seed=0, tiers=4, samples_per_tier=5)
rbgs = bench.build_fleet(1024, n)
blob, _ = B200TopoPodGroupManager(OraclePlacer(topo)).groups_blob(rbgs)
blob = np.ascontiguousarray(blob, dtype=np.int32)
degp1 = (np.diff(topo.row_ptr) + 1).astype(np.int32)
wsum = int(max(topo.edge_w[topo.row_ptr[i]:topo.row_ptr[i + 1]].sum() for i in range(n)))
lib = _lib.load()
I32P = C.POINTER(C.c_int32)
ns, nw, pw = C.c_int32(), C.c_int32(), C.c_int64()
args = (blob.ctypes.data_as(I32P), len(blob), n, int(topo.domain.max()) + 1, degp1.ctypes.data_as(I32P), wsum,
        None, 0, C.byref(ns), C.byref(nw), C.byref(pw))
for _ in range(20):
    assert lib.rbgtopo_plan_describe(*args) == 0
reps = 300
best = 1e9
for _ in range(5):
    t0 = time.perf_counter()
    for _ in range(reps):
        lib.rbgtopo_plan_describe(*args)
    best = min(best, (time.perf_counter() - t0) / reps * 1e6)
print(f"groups_blob {len(blob)} words, {ns.value} steps, {nw.value} waves, plan {pw.value} words: {best:.1f} us per call")
