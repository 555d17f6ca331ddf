"""CPU: the fast CPU variant (oracle/placer_fast.c: the GPU path's algebra on the host — base vector
per snapshot, sparse patches, partial selection) against the literal oracle (oracle/placer_oracle.c):
dense matrix bits, top-K keys, assignment, status, domain on random step batches and whole fleets."""
import numpy as np
import pytest

import bench
from oracle import placer as oracle_placer
from oracle import wave_loop
from rbg_b200 import synth
from test_gpu_parity import _random_steps


def _same(a, b):
    assert a["rc"] == b["rc"] == 0
    assert np.array_equal(a["matrix"].view(np.uint32), b["matrix"].view(np.uint32))
    assert np.array_equal(a["topk"], b["topk"])
    assert np.array_equal(a["assign"], b["assign"]) and np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["domain"], b["domain"])


@pytest.mark.parametrize("n,seed,excl,gang", [(512, 1, False, False), (3000, 2, True, True), (4096, 3, True, False),
                                              (700, 4, False, True)])
def test_fast_equals_literal_on_random_steps(n, seed, excl, gang):
    topo = synth.make_topology(n, seed=seed, tiers=4, owned_frac=0.25 if excl else 0.0)
    for k in range(3):
        blob = _random_steps(topo, 10 * seed + k, 24, excl=excl, gang=gang)
        _same(oracle_placer.place(topo, blob), oracle_placer.place_fast(topo, blob))
        _same(oracle_placer.place(topo, blob), oracle_placer.place_fast(topo, blob, nthreads=4))


def test_fast_equals_literal_on_scarce_capacity_and_changing_snapshots():
    topo = synth.make_topology(600, seed=7, tiers=3, max_free=2)
    blob = _random_steps(topo, 5, 30, excl=False, gang=True)
    _same(oracle_placer.place(topo, blob), oracle_placer.place_fast(topo, blob))
    topo.free = np.where(np.arange(600) % 3 == 0, 0, topo.free).astype(np.int32)   # new snapshot, same arrays' shape
    _same(oracle_placer.place(topo, blob), oracle_placer.place_fast(topo, blob))


def test_fast_wave_loop_of_a_fleet():
    topo = synth.make_topology(2000, seed=11, tiers=4)
    specs = bench.fleet_spec("mooncake", 16, 2000, seed=3)
    _, blobs = wave_loop.run_fleet(topo, bench.to_oracle(specs))
    for b in blobs:
        _same(oracle_placer.place(topo, b), oracle_placer.place_fast(topo, b, nthreads=2))
