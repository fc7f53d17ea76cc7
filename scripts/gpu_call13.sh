#!/bin/bash
# Round-2 session L: loads issued before the early exits (load_now) vs the compiler's placement, A/B.
tag=${1:-r2t}
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
P=$PWD/deodr_b200/libdeodr_b200_plain.so
run_bench c5_now c5 40 X=1
run_bench c5_plain c5 40 DEODR_B200_LIB=$P
run_bench c5_now2 c5 40 X=1
run_bench c5_plain2 c5 40 DEODR_B200_LIB=$P
run_bench c2_now c2 200 X=1
run_bench c2_plain c2 200 DEODR_B200_LIB=$P
run_bench c4_now c4 20 X=1
run_bench c4_plain c4 20 DEODR_B200_LIB=$P
run_bench c3_now c3 50 X=1
run_bench c3_plain c3 50 DEODR_B200_LIB=$P
