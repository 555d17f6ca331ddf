"""Multi-GPU node-axis sharding through the C ABI (rbgtopo_shard_*), one process
per GPU over NCCL.  Needs >= 2 GPUs; launched by the test with torch.distributed.run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["RBG_ROOT"]); sys.path.insert(0, os.path.join(os.environ["RBG_ROOT"], "tests"))
from oracle import placer as oracle_placer
from rbg_b200 import synth
from rbg_b200.engine import TopoPlacer
from test_gpu_parity import _random_steps
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
class DevPtr:
    def __init__(self, p, nb):
        self.__cuda_array_interface__ = {"shape": (nb // 8,), "typestr": "<i8", "data": (p, False), "version": 3, "strides": None}
for n, seed, excl in [(4096, 1, False), (10000, 2, True), (3000, 3, True)]:
    topo = synth.make_topology(n, seed=seed, tiers=4, owned_frac=0.25 if excl else 0.0)
    blob = _random_steps(topo, 50 + seed, 24, excl=excl, gang=True)
    ref = oracle_placer.place(topo, blob)
    eng = TopoPlacer(device=local, rank=rank, world=world)
    eng.set_topology(topo.row_ptr, topo.col_idx, topo.edge_w, topo.free, topo.domain, topo.domain_owner)
    stream = torch.cuda.Stream()                 # kernels and NCCL ordered on one stream
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    h = eng.stage(blob)
    p, nb = eng.shard_score(h)
    src = torch.as_tensor(DevPtr(p, nb), device="cuda")
    allk = torch.empty(world * (nb // 8), dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(allk, src)
    need2, p2, nb2 = eng.shard_merge(h, allk.data_ptr())
    all2 = None
    if need2:
        src2 = torch.as_tensor(DevPtr(p2, nb2), device="cuda")
        all2 = torch.empty(world * (nb2 // 8), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(all2, src2)
    eng.shard_assign(h, all2.data_ptr() if all2 is not None else None)
    assign, status, domain = eng.fetch(h)
    assert np.array_equal(assign, ref["assign"]), (rank, n, assign[:16], ref["assign"][:16])
    assert np.array_equal(status, ref["status"]) and np.array_equal(domain, ref["domain"])
    lo, hi = eng.slab()
    for row in range(0, ref["matrix"].shape[0], 7):          # this rank's column slab of the dense matrix
        got = eng.read_scores(h, row)
        assert np.array_equal(got.view(np.uint32), ref["matrix"][row, lo:hi].view(np.uint32)), (rank, row)
    for rr in range(ref["topk"].shape[0]):
        assert np.array_equal(eng.read_topk(h, rr, 32), ref["topk"][rr]), (rank, rr)
    eng.release(h)
    # the same batch with replicated selection: no collective, plain run_staged / score_assign on every rank
    h = eng.stage(blob)
    eng.run_staged(h, 1)
    a2, s2, d2 = eng.fetch(h)
    assert np.array_equal(a2, ref["assign"]) and np.array_equal(s2, ref["status"]) and np.array_equal(d2, ref["domain"]), (rank, n)
    for row in range(0, ref["matrix"].shape[0], 5):
        got = eng.read_scores(h, row)
        assert np.array_equal(got.view(np.uint32), ref["matrix"][row, lo:hi].view(np.uint32)), (rank, row, "replicated")
    for rr in range(ref["topk"].shape[0]):
        assert np.array_equal(eng.read_topk(h, rr, 32), ref["topk"][rr]), (rank, rr, "replicated")
    eng.release(h)
    a3, s3, d3 = eng.score_assign(blob)
    assert np.array_equal(a3, ref["assign"]) and np.array_equal(s3, ref["status"]), (rank, n)
    eng.close()
# ---- whole groups through the sharded multi-wave plan (one emit launch per rank, per-wave
#      select -> all-gather -> merge -> assign, placements chained on every rank)
from rbg_b200.plugin import B200TopoPodGroupManager
from test_gpu_groups import _fleet
from test_plugin_host import OraclePlacer
for n, kw in [(8000, {}), (6000, dict(excl_every=3, gang_every=4)), (5000, dict(big_every=5))]:
    topo = synth.make_topology(n, seed=n, tiers=4, owned_frac=0.2 if kw.get("excl_every") else 0.0)
    rbgs = _fleet(n, 24, seed=9, **kw)
    ref = B200TopoPodGroupManager(OraclePlacer(topo)).reconcile_pod_groups_by_waves(rbgs)
    eng = TopoPlacer(device=local, rank=rank, world=world)
    eng.set_topology(topo.row_ptr, topo.col_idx, topo.edge_w, topo.free, topo.domain, topo.domain_owner)
    eng.set_stream(stream.cuda_stream)
    gblob, runs = B200TopoPodGroupManager(eng).groups_blob(rbgs)
    h = eng.stage_groups(gblob)
    for w in range(eng.shard_waves(h)):
        p, nb = eng.shard_wave_score(h, w)
        src = torch.as_tensor(DevPtr(p, nb), device="cuda")
        allk = torch.empty(world * (nb // 8), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allk, src)
        need2, p2, nb2 = eng.shard_wave_merge(h, w, allk.data_ptr())
        all2 = None
        if need2:
            src2 = torch.as_tensor(DevPtr(p2, nb2), device="cuda")
            all2 = torch.empty(world * (nb2 // 8), dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(all2, src2)
        eng.shard_wave_assign(h, w, all2.data_ptr() if all2 is not None else None)
    assign, status, domain = eng.fetch(h)
    off = 0
    for i, (g, r) in enumerate(zip(runs, ref)):
        want = list(r.nodes.values())
        got = assign[off:off + len(want)].tolist()
        off += len(want)
        if r.status == 1:      # plan leaves non-gang partial groups to the host loop: only the status is checked
            assert status[i] == 1, (rank, n, i)
            continue
        assert got == want, (rank, n, i, got, want)
        assert status[i] == r.status and domain[i] == r.domain, (rank, n, i, status[i], r.status, domain[i], r.domain)
    eng.release(h)
    # ---- the same plan with REPLICATED selection (k_plan_group over all nodes on every rank, the
    #      dense matrix still column-sharded): no collective at all, identical placements on every rank
    h = eng.stage_groups(gblob)
    eng.run_staged(h, 1)
    assign2, status2, domain2 = eng.fetch(h)
    assert np.array_equal(assign2, assign) and np.array_equal(status2, status) and np.array_equal(domain2, domain), (rank, n)
    a4, s4, d4 = eng.place_groups(gblob)     # plan + exact host loop for partially placed groups
    off = 0
    for i, r in enumerate(ref):
        want = list(r.nodes.values())
        assert a4[off:off + len(want)].tolist() == want and s4[i] == r.status and d4[i] == r.domain, (rank, n, i)
        off += len(want)
    if not kw:      # nobody fails in this fleet: the plan's rows line up with the wave-by-wave oracle run
        from oracle import placer as oracle_placer
        from rbg_b200.blob import BlobBuilder
        from rbg_b200.plugin import _GroupRun
        from gpu_util import plan_rows
        gruns = [_GroupRun(r, B200TopoPodGroupManager(eng).arith) for r in rbgs]
        row_of = plan_rows(gblob, topo)
        index_of = {id(g): i for i, g in enumerate(gruns)}
        lo, hi = eng.slab()
        w = 0
        while True:
            active = [g for g in gruns if w < len(g.waves)]
            if not active:
                break
            bb = BlobBuilder()
            for g in active:
                bb.add(g.step(w))
            oref = oracle_placer.place(topo, bb.build(), want_matrix=True, want_topk=False)
            assert oref["rc"] == 0 and (oref["status"] == 0).all()
            off = 0
            for i, g in enumerate(active):
                cnt = sum(c for _, _, c in g.waves[w].roles)
                row0 = row_of[(index_of[id(g)], w)]
                for k in range(0, cnt, 2):
                    got = eng.read_scores(h, row0 + k)
                    assert np.array_equal(got.view(np.uint32), oref["matrix"][off + k, lo:hi].view(np.uint32)), (rank, w, i, k)
                g.absorb(w, oref["assign"][off:off + cnt], int(oref["status"][i]), int(oref["domain"][i]), n)
                off += cnt
            w += 1
    eng.release(h); eng.close()
dist.barrier()
if rank == 0: print("SHARD_OK", world)
dist.destroy_process_group()
'''


def test_node_axis_sharding_matches_oracle(tmp_path):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        # One GPU: NCCL needs one device per rank, so the same protocol runs with both ranks as contexts of this
        # device and a device copy as the all-gather (tests/test_gpu_shard_single.py) — same library entry points,
        # same oracle checks; the NCCL transport itself is exercised on the >= 2-GPU boxes (profiles/README.md).
        import test_gpu_shard_single as single
        single.test_step_batches_sharded_on_one_device(2)
        single.test_group_plans_sharded_on_one_device(2)
        return
    world = 2 if ngpu < 4 else 4
    script = tmp_path / "shard_worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, RBG_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARD_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
