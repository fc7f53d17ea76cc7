timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
for i in 1 2; do
for v in old new; do
DEODR_B200_LIB=build/ab/libdeodr_$v.so timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/r33_$v$i.json
python -c "
import json;d=json.load(open('gpurun_out/r33_$v$i.json'));print('$v',d['ms_per_step'],d['roofline']['phase_ms'])"
done; done
for v in old new; do
DEODR_B200_SERIAL=1 DEODR_B200_LIB=build/ab/libdeodr_$v.so timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/r33_serial_$v.json
python -c "
import json;d=json.load(open('gpurun_out/r33_serial_$v.json'));print('serial $v',d['ms_per_step'],d['roofline']['phase_ms'])"
done
