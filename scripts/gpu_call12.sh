#!/bin/bash
# Round-2 session K: record-parallel adjoint of the small triangles (k_small_rec_bwd) - validation + A/B.
tag=${1:-r2s}
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
python -m pytest tests/test_gpu_parity.py tests/test_gpu_views.py tests/test_gpu_api.py -q -x -p no:cacheprovider --timeout 600 > gpurun_out/${tag}_pytest_a.log 2>&1; echo "pytest (record) rc=$?"; tail -3 gpurun_out/${tag}_pytest_a.log
DEODR_B200_SMALL_ADJOINT=triangle python -m pytest tests/test_gpu_parity.py -k "golden or soup or micro or meshes or config3" -q -x -p no:cacheprovider --timeout 600 > gpurun_out/${tag}_pytest_b.log 2>&1; echo "pytest (triangle) rc=$?"; tail -2 gpurun_out/${tag}_pytest_b.log
run_bench c5_rec c5 40 X=1
run_bench c5_tri c5 40 DEODR_B200_SMALL_ADJOINT=triangle
run_bench c5_rec2 c5 40 X=1
run_bench c4_rec c4 20 X=1
run_bench c4_tri c4 20 DEODR_B200_SMALL_ADJOINT=triangle
run_bench c5t_rec c5t 30 X=1
run_bench c5t_tri c5t 30 DEODR_B200_SMALL_ADJOINT=triangle
timeout 200 compute-sanitizer --tool racecheck python scripts/sanitize_scenes.py > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -1 gpurun_out/${tag}_racecheck.log
timeout 200 compute-sanitizer --tool memcheck python scripts/sanitize_scenes.py > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -1 gpurun_out/${tag}_memcheck.log
