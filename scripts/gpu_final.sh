#!/bin/bash
# End-of-round session: the -m gpu suite, smoke, the bench lines of every workload, the reference arm, then the ncu
# launch list + one --set full capture of the raster kernels on the headline workload.  Everything lands in gpurun_out/.
tag=${1:-r2f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
# ncu first: the bench lines that follow read roofline.traffic from the capture of THIS build
bash scripts/gpu_profile.sh ${tag} c5
python scripts/make_profile_summary.py ${tag}_c5 gpurun_out/launches_${tag}.csv gpurun_out/prof_${tag}.ncu-rep > gpurun_out/${tag}_summary.log 2>&1
cp profiles/ncu_traffic.json gpurun_out/${tag}_ncu_traffic.json
python bench.py > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
python bench.py --workload c4 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
python bench.py --workload c2 --steps 200 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
python bench.py --workload c3 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err
for f in c5 c4 c2 c3; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_$f.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["config"]["timed_region"], "eager", d["config"]["eager_ms_per_step"], "launches", d["gpu_launches"], "e2e", (d.get("e2e") or {}).get("ms_per_step"))
    print(d["roofline"]["phase_ms"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/${tag}_bench_$f.err").read()[-800:])
PY
done
tail -c 400 gpurun_out/${tag}_bench_ref.json
