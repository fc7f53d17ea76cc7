#!/bin/bash
# ncu session on the headline workload: serialised launch list (shares) + one --set full capture of the raster kernels,
# then compute-sanitizer over scripts/sanitize_scenes.py.
tag=${1:-r2}
wl=${2:-c5}
mkdir -p gpurun_out
export DEODR_B200_SERIAL=1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --eager > gpurun_out/ncu_${tag}_1.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k 'regex:^(k_bin|k_tile_z|k_shade|k_edge_fwd|k_raster_bwd|k_small_tri_bwd|k_interior_bwd|k_bin_edges|k_sort_tile_edges|k_finalize_edges)$' -s 27 -c 9 \
    -f -o gpurun_out/prof_${tag} python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --eager > gpurun_out/ncu_${tag}_2.log 2>&1
ls -la gpurun_out/prof_${tag}.ncu-rep gpurun_out/launches_${tag}.csv
unset DEODR_B200_SERIAL
if [ -n "$SANITIZE" ]; then
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python scripts/sanitize_scenes.py > gpurun_out/sanitizer_${tag}_$tool.log 2>&1
  tail -3 gpurun_out/sanitizer_${tag}_$tool.log
done
fi
