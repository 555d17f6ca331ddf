"""2+ GPU probe of the in-library all-gather (rbgtopo_run_staged_p2p): one process per GPU under
torchrun, verbose, short timeouts.  usage: torchrun --nproc-per-node N profiles/probes/p2p_probe.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import placer as oracle_placer
from rbg_b200 import synth
from rbg_b200.engine import TopoPlacer
from test_gpu_parity import _random_steps

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def say(*a):
    print(f"[rank {rank}]", *a, flush=True)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
class D: pass
D.world, D.dist = world, dist
D.barrier = lambda: (dist.barrier(), torch.cuda.synchronize())
topo = synth.make_topology(4096, seed=1, tiers=4)
blob = _random_steps(topo, 71, 24, excl=False, gang=True)
ref = oracle_placer.place(topo, blob)
eng = TopoPlacer(device=local, rank=rank, world=world)
eng.set_topology(topo.row_ptr, topo.col_idx, topo.edge_w, topo.free, topo.domain, topo.domain_owner)
say("peer access:", [torch.cuda.can_device_access_peer(local, g) for g in range(world) if g != local])
eng.p2p_connect(D)
say("connected")
h = eng.stage(blob)
t0 = time.time()
eng.run_staged_p2p(h, 1)
say("enqueued")
torch.cuda.synchronize()
say("synced after %.3f s" % (time.time() - t0), eng.p2p_stats())
a, s, d = eng.fetch(h)
say("assign ok:", np.array_equal(a, ref["assign"]), "status ok:", np.array_equal(s, ref["status"]))
for it in range(3):
    eng.run_staged_p2p(h, 1)
torch.cuda.synchronize()
say("3 more passes", eng.p2p_stats())
a, s, d = eng.fetch(h)
say("assign ok:", np.array_equal(a, ref["assign"]))
t0 = time.time()
eng.run_staged_p2p(h, 50)
torch.cuda.synchronize()
say("50 passes in %.2f ms" % ((time.time() - t0) * 1e3), eng.p2p_stats())
eng.release(h); eng.close()
dist.barrier()
dist.destroy_process_group()
say("done")
