#!/bin/bash
tag=${1:-r2r}
mkdir -p gpurun_out
run_bench() {  # name workload steps env...
  local name=$1 wl=$2 steps=$3; shift 3
  env "$@" python bench.py --workload $wl --steps $steps --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_${name}.json").read().strip().splitlines()[-1])
    print("${name}:", d["ms_per_step"], "eager", d["config"]["eager_ms_per_step"], d["roofline"].get("phase_ms"))
except Exception as e:
    print("${name}: FAILED", e); print(open("gpurun_out/${tag}_${name}.err").read()[-600:])
PY
}
run_bench c4_lanes4 c4 20 X=1
run_bench c4_lanes6 c4 20 DEODR_B200_LANES=6
run_bench c4_lanes8 c4 20 DEODR_B200_LANES=8
python -m pytest tests/test_gpu_views.py -k "config4 or batch" -q -x -p no:cacheprovider --timeout 600 2>&1 | tail -2
