#!/bin/bash
# One GPU session: the -m gpu suite, then the bench lines (c5 default, c4, c2 eager + graph).  Everything lands in gpurun_out/.
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err
python bench.py --workload c4 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
python bench.py --workload c2 --steps 200 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
python bench.py --workload c2 --steps 200 --warmup 10 --no-e2e --no-cpu-baseline --graph > gpurun_out/${tag}_bench_c2_graph.json 2> gpurun_out/${tag}_bench_c2_graph.err
python bench.py --workload c3 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err
for f in c5 c4 c2 c2_graph c3; do echo "== $f"; cut -c1-400 gpurun_out/${tag}_bench_$f.json; tail -2 gpurun_out/${tag}_bench_$f.err; done
