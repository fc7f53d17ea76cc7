#!/bin/bash
# quick multi-workload comparison: ms/step of c5 c3 c2 c4 for the given environment spec (one line each)
tag=$1; shift
for wl in c5 c3 c2 c4; do
  steps=40; [ $wl = c2 ] && steps=200
  env "$@" python bench.py --workload $wl --steps $steps --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_$wl.json 2> gpurun_out/${tag}_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_$wl.json").read().strip().splitlines()[-1])
    print("$tag $wl:", d["ms_per_step"], d["roofline"].get("phase_ms"))
except Exception as e:
    print("$tag $wl FAILED", e); print(open("gpurun_out/${tag}_$wl.err").read()[-600:])
PY
done
