#!/bin/bash
# Multi-GPU evidence on one box: the NCCL shard test, then bench.py per shard mode.
# usage (under gpurun --gpus N): bash profiles/probes/r2_multi.sh N TAG [modes...]
N=$1; tag=$2; shift; shift
modes=${@:-replicated p2p allgather}
python -m pytest tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_shardtest_n$N.log
port=29600
for m in $modes; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --shard-mode $m --no-cpu > gpurun_out/${tag}_n${N}_$m.json 2> gpurun_out/${tag}_n${N}_$m.err || tail -8 gpurun_out/${tag}_n${N}_$m.err
  python - "$m" "gpurun_out/${tag}_n${N}_$m.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    a = d.get("alt", {})
    print(sys.argv[1], "cfg3: %.3e scores/s, %.4f ms/step, e2e %.3e (%.4f ms), parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity"]["ok"]))
    for k, v in a.items():
        if "value" in v:
            print("   ", k, "%.3e scores/s, %.4f ms/step, e2e %.3e (%.4f ms), parity %s" % (v["value"], v["ms_per_step"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v["parity"]["ok"]))
        else:
            print("   ", k, "e2e %.3e (%.4f ms), parity %s" % (v["e2e"]["value"], v["e2e"]["ms_per_step"], v["parity"]["ok"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
