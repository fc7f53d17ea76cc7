#!/bin/bash
# Round-2 session J: warp-per-edge k_bin_edges (validation + small-scene bench lines), lanes A/B on c4.
tag=${1:-r2q}
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
python -m pytest tests/test_gpu_parity.py tests/test_gpu_views.py -k "not config5 and not config4" -q -x -p no:cacheprovider --timeout 600 > gpurun_out/${tag}_pytest_a.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_a.log
run_bench c2 c2 200 X=1
run_bench c5 c5 40 X=1
run_bench c3 c3 60 X=1
run_bench c4 c4 20 X=1
run_bench c4_lanes3 c4 20 DEODR_B200_LANES=3
run_bench c4_lanes4 c4 20 DEODR_B200_LANES=4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
