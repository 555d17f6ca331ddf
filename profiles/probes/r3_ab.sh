#!/bin/bash
# A/B of library builds / switches on one box: bench.py value leg per variant.
# usage (under gpurun): bash profiles/probes/r3_ab.sh TAG 'label:ENV=VAL,ENV=VAL' ...   (label alone = defaults)
tag=$1; shift
for spec in "$@"; do
  label=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  env $envs python bench.py --no-cpu --no-alt > gpurun_out/${tag}_$label.json 2> gpurun_out/${tag}_$label.err || tail -5 gpurun_out/${tag}_$label.err
  python - "$label" "gpurun_out/${tag}_$label.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    r = d["roofline"]
    print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "(kernel-timing leg %.4f)" % d.get("ms_per_step_kernel_timing", 0), "emit med %.1f" % r["kernel_launch_us"]["median"],
          "select med %.1f" % r["select_launch_us"]["median"], "e2e ms %.4f" % d["e2e"]["ms_per_step"], "parity", d["parity"]["ok"],
          "frac %.3f step %.3f" % (r["frac"], r["whole_step_frac"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
