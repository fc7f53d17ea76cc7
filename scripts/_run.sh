set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2> gpurun_out/r21_new_$i.err | tee gpurun_out/r21_new_$i.json | cut -c1-300
grep -i "phase\|tile_z" gpurun_out/r21_new_$i.err | tail -3
DEODR_B200_TILEZ_CTAS_PER_SM=0 DEODR_B200_LIB=build/ab/libdeodr_old.so timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2> gpurun_out/r21_old_$i.err | tee gpurun_out/r21_old_$i.json | cut -c1-300
grep -i "phase\|tile_z" gpurun_out/r21_old_$i.err | tail -3
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --workload c3 2>gpurun_out/r21_c3.err | cut -c1-400
tail -3 gpurun_out/r21_c3.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --workload c2 2>gpurun_out/r21_c2.err | cut -c1-400
tail -3 gpurun_out/r21_c2.err
