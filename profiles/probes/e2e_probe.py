import os, sys, time, numpy as np
os.environ["RBGTOPO_PROFILE_HOST"] = "1"
sys.path.insert(0, '/root/repo')
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
for _ in range(3):
    eng.update_nodes(free); eng.place_groups(gblob)
tu = tp = 0
for _ in range(20):
    t0 = time.perf_counter(); eng.update_nodes(free); t1 = time.perf_counter(); eng.place_groups(gblob); t2 = time.perf_counter()
    tu += t1 - t0; tp += t2 - t1
print("update_nodes ms", tu / 20 * 1e3, "place_groups ms", tp / 20 * 1e3, eng.last_timing())
