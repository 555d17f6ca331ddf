#!/bin/bash
# A/B of the plan pipeline variants on one box: bench.py value leg per launch-order / priority switch.
# usage (under gpurun): bash profiles/probes/r2_ab.sh TAG VAR1 VAR2 ...   (VAR = env switch set to 1, DEFAULT = none)
tag=$1; shift
for v in "$@"; do
  envs=$(echo "$v" | tr '+' '\n' | sed 's/$/=1/' | tr '\n' ' ')   # A+B -> A=1 B=1;  X=3 style values: write X:3
  envs=$(echo "$envs" | sed 's/:\([0-9]*\)=1/=\1/g')
  env $envs python bench.py --no-cpu --no-alt > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err || tail -5 gpurun_out/${tag}_$v.err
  python - "$v" "gpurun_out/${tag}_$v.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    r = d["roofline"]
    print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "emit", r["kernel_launch_us"], "select", r["select_launch_us"],
          "e2e ms %.4f" % d["e2e"]["ms_per_step"], "parity", d["parity"]["ok"], "frac %.3f step %.3f" % (r["frac"], r["whole_step_frac"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
