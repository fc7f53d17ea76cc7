timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
for i in 1 2; do
for v in old new small6 small7 int4 int3 shade5 shade8 edge2 edge4; do
DEODR_B200_SERIAL=1 DEODR_B200_LIB=build/ab/libdeodr_$v.so timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/r39_$v$i.json
python -c "
import json;d=json.load(open('gpurun_out/r39_$v$i.json'));p=d['roofline']['phase_ms'];print('$v',d['ms_per_step'],'shade',p['shade'],'edge_fwd',p['edge_fwd'],'edge_bwd',p['edge_bwd'],'int',p['interior_bwd'],'small',p['small_tri_bwd'])"
done; done
