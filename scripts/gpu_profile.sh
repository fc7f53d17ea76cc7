#!/bin/bash
# ncu session on the headline workload: serialised launch list (shares) + one --set full capture of the raster kernels.
tag=${1:-r2}
wl=${2:-c5}
mkdir -p gpurun_out
export DEODR_B200_SERIAL=1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_${tag}_1.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k 'regex:^(k_bin|k_tile_z|k_shade|k_edge_fwd|k_raster_bwd|k_small_tri_bwd|k_bin_edges|k_sort_tile_edges)$' -s 25 -c 8 \
    -f -o gpurun_out/prof_${tag} python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_${tag}_2.log 2>&1
tail -3 gpurun_out/ncu_${tag}_1.log gpurun_out/ncu_${tag}_2.log
ls -la gpurun_out/prof_${tag}.ncu-rep gpurun_out/launches_${tag}.csv
