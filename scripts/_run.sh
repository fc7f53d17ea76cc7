timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
python scripts/gpu_parity_quick.py 2>&1 | grep -c "exact=True"
for i in 1 2; do
for v in old new tz5 tz3; do
DEODR_B200_SERIAL=1 DEODR_B200_LIB=build/ab/libdeodr_$v.so timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/r40_$v$i.json
python -c "
import json;d=json.load(open('gpurun_out/r40_$v$i.json'));p=d['roofline']['phase_ms'];print('$v',d['ms_per_step'],'tile_z',p['tile_z'],'shade',p['shade'],'edge_fwd',p['edge_fwd'],'edge_bwd',p['edge_bwd'],'int',p['interior_bwd'],'small',p['small_tri_bwd'])"
done; done
