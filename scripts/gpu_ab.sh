#!/bin/bash
# A/B of library variants / environment switches on the headline workload; each arm twice, interleaved.
tag=${1:-ab}; shift
wl=${WL:-c5}
mkdir -p gpurun_out
for rep in 1 2; do
  for arm in "$@"; do
    name=${arm%%:*}; spec=${arm#*:}
    env $spec python bench.py --workload $wl --steps 40 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_${name}_$rep.json 2> gpurun_out/${tag}_${name}_$rep.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_${name}_$rep.json").read().strip().splitlines()[-1])
    print("${name} rep $rep:", d["ms_per_step"], d["roofline"].get("phase_ms"))
except Exception as e:
    print("${name} rep $rep: FAILED", e); print(open("gpurun_out/${tag}_${name}_$rep.err").read()[-800:])
PY
  done
done
