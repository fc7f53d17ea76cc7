#!/bin/bash
# Round-2 session C: occupancy variants of the main kernels + the small_textured threshold on a 1M-triangle textured scene.
tag=${1:-r2j}
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
python -m pytest tests/test_gpu_parity.py -k "golden or soup or ties or meshes or more_prim or micro" -q -x -p no:cacheprovider --timeout 600 > gpurun_out/${tag}_pytest_a.log 2>&1; echo "pytest default rc=$?"; tail -3 gpurun_out/${tag}_pytest_a.log
A=$PWD/deodr_b200/libdeodr_b200_vA.so; B=$PWD/deodr_b200/libdeodr_b200_vB.so
run_bench c5_base c5 40 X=1
run_bench c5_vA c5 40 DEODR_B200_LIB=$A
run_bench c5_vB c5 40 DEODR_B200_LIB=$B
run_bench c5t_st1 c5t 30 DEODR_B200_SMALL_TEXTURED=1
run_bench c5t_st0 c5t 30 DEODR_B200_SMALL_TEXTURED=0
run_bench c3_base c3 60 DEODR_B200_SMALL_TEXTURED=0
run_bench c3_vA c3 60 DEODR_B200_SMALL_TEXTURED=0 DEODR_B200_LIB=$A
run_bench c3_vB c3 60 DEODR_B200_SMALL_TEXTURED=0 DEODR_B200_LIB=$B
run_bench c2_base c2 200 X=1
run_bench c2_vA c2 200 DEODR_B200_LIB=$A
