#!/bin/bash
# Round-2 session B: validation of the chunk-pipelined z pass + the small_textured switch, A/B bench arms, racecheck, e2e trace.
tag=${1:-r2i}
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
python -m pytest tests/test_gpu_parity.py tests/test_gpu_views.py -k "not config5" -q -x -p no:cacheprovider --timeout 600 > gpurun_out/${tag}_pytest_a.log 2>&1; echo "pytest default rc=$?"; tail -3 gpurun_out/${tag}_pytest_a.log
DEODR_B200_SMALL_TEXTURED=0 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -k "golden or soup or config3 or more_prim or ties or autograd or scene2d" -q -x -p no:cacheprovider --timeout 600 > gpurun_out/${tag}_pytest_b.log 2>&1; echo "pytest small_textured=0 rc=$?"; tail -3 gpurun_out/${tag}_pytest_b.log
run_bench c5_a c5 40 A=1
run_bench c5_tilez4 c5 40 DEODR_B200_LIB=$PWD/deodr_b200/libdeodr_b200_tilez4.so
run_bench c5_tilez6 c5 40 DEODR_B200_LIB=$PWD/deodr_b200/libdeodr_b200_tilez6.so
run_bench c5_b c5 40 A=1
run_bench c3_smalltex1 c3 60 DEODR_B200_SMALL_TEXTURED=1
run_bench c3_smalltex0 c3 60 DEODR_B200_SMALL_TEXTURED=0
run_bench c4 c4 20 A=1
run_bench c2 c2 200 A=1
timeout 200 compute-sanitizer --tool racecheck python scripts/sanitize_scenes.py > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${tag}_racecheck.log
python scripts/e2e_trace.py c5 > gpurun_out/${tag}_e2e_trace.log 2>&1; tail -14 gpurun_out/${tag}_e2e_trace.log
