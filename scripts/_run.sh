timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
python scripts/gpu_parity_quick.py 2>&1 | grep -c "exact=True"
python scripts/gpu_parity_quick.py 2>&1 | grep -A1 "torus158\|torus40" | head -6
for i in 1 2; do
for v in new tex notex; do
DEODR_B200_LIB=build/ab/libdeodr_$v.so timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$v overlapped',d['ms_per_step'])"; done; done
for w in c3 c2; do for v in new tex; do DEODR_B200_LIB=build/ab/libdeodr_$v.so timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --workload $w 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$w $v',d['ms_per_step'])"; done; done
