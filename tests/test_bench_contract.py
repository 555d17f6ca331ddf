"""CPU: the parts of bench.py's contract that do not need a GPU — the reference arm
(the CPU oracle timed on the host cores) prints one JSON line with the agreed keys, rank != 0
stays silent under torchrun, and the default arguments match the driver's invocation."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *args):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=300)


def test_reference_arm_prints_the_contract_line():
    r = _run(None, "--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-groups", "8", "--groups", "16",
             "--nodes", "1000")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "replica_x_node_affinity_scores_per_sec"
    assert d["unit"] == "scores/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_silently():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--steps", "1",
             "--warmup", "0", "--ref-groups", "4", "--groups", "8", "--nodes", "500")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_default_arguments():
    sys.path.insert(0, ROOT)
    import bench
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert '"--gpus", type=int, default=1' in src
    assert bench.METRIC == "replica_x_node_affinity_scores_per_sec" and bench.UNIT == "scores/s"
