"""Host / device split of the end-to-end call (rbgtopo_update_nodes + rbgtopo_place_groups with host
buffers) on the bench fleet.  RBGTOPO_PROFILE_HOST=1 prints the host phases of every call on stderr;
PROBE_QUIET=1 times without them.  Environment switches of the library apply (RBGTOPO_NO_EARLY_EMIT ...)."""
import os, sys, time, numpy as np
if not os.environ.get("PROBE_QUIET"):
    os.environ["RBGTOPO_PROFILE_HOST"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from rbg_b200 import synth
from rbg_b200.engine import TopoPlacer
from rbg_b200.plugin import B200TopoPodGroupManager
topo = synth.make_topology(10000, seed=0, tiers=4, samples_per_tier=5)
rbgs = bench.build_fleet(1024, 10000)
eng = TopoPlacer(device=0)
eng.set_topology(topo.row_ptr, topo.col_idx, topo.edge_w, topo.free, topo.domain, topo.domain_owner)
gblob, _ = B200TopoPodGroupManager(eng).groups_blob(rbgs)
free = np.ascontiguousarray(topo.free, dtype=np.int32)
for _ in range(30):
    eng.update_nodes(free); eng.place_groups(gblob)
best = (1e9, 0, 0)
for rnd in range(5):
    tu = tp = 0
    for _ in range(20):
        t0 = time.perf_counter(); eng.update_nodes(free); t1 = time.perf_counter(); eng.place_groups(gblob); t2 = time.perf_counter()
        tu += t1 - t0; tp += t2 - t1
    if tu + tp < best[0]:
        best = (tu + tp, tu, tp)
print("best of 5 rounds: update_nodes ms %.4f place_groups ms %.4f total %.4f" % (best[1] / 20 * 1e3, best[2] / 20 * 1e3, best[0] / 20 * 1e3), eng.last_timing())
