timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
for i in 1 2 3; do
for v in old new; do
DEODR_B200_TRACE_GAPS=1 DEODR_B200_LIB=build/ab/libdeodr_$v.so timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>gpurun_out/r37.err > gpurun_out/r37_$v$i.json
python -c "
import json;d=json.load(open('gpurun_out/r37_$v$i.json'));p=d['roofline']['phase_ms'];print('$v',d['ms_per_step'],'bin_count',p['bin_count'])"
grep "gap before edge_order\|gap before bin_fill" gpurun_out/r37.err
done; done
timeout 300 python scripts/e2e_trace.py c5 2>&1 | grep "iter 3"
