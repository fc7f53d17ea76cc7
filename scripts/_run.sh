DEODR_B200_SERIAL=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/launches_r1n.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_n1.log 2>&1
DEODR_B200_SERIAL=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_ -s 56 -c 14 -f -o gpurun_out/prof_r1n python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_n2.log 2>&1
tail -n 1 gpurun_out/ncu_n2.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r42_bench.json 2> gpurun_out/r42_bench.err; cut -c1-300 gpurun_out/r42_bench.json
for w in c3 c2; do timeout 300 python bench.py --steps 30 --warmup 5 --workload $w > gpurun_out/r42_$w.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r42_$w.json'));print('$w',d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['e2e']['value'],d['cpu_baseline']['value'])"; done
