#!/bin/bash
for arm in "base:X=1" "notma:DEODR_B200_TMA_TILES=0" "nofuse:DEODR_B200_FUSE_SHADE=0" "serial:DEODR_B200_SERIAL=1" "rows16:DEODR_B200_RECORD_ROWS=16"; do
  name=${arm%%:*}; spec=${arm#*:}
  env $spec python bench.py --workload c2 --steps 300 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/c2ab_$name.json 2> gpurun_out/c2ab_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/c2ab_$name.json").read().strip().splitlines()[-1])
print("$name", d["ms_per_step"], "eager", d["config"]["eager_ms_per_step"], d["roofline"]["phase_ms"])
PY
done
