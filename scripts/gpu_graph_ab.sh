#!/bin/bash
tag=$1
for wl in c3 c3u c2 c5; do
 for arm in "fuse:X=1" "nofuse:DEODR_B200_FUSE_SHADE=0"; do
  name=${arm%%:*}; spec=${arm#*:}
  for mode in "" "--graph"; do
   steps=100; [ $wl = c2 ] && steps=300
   env $spec python bench.py --workload $wl --steps $steps --warmup 5 --no-e2e --no-cpu-baseline $mode > gpurun_out/${tag}_${wl}_${name}${mode}.json 2> gpurun_out/${tag}_${wl}_${name}${mode}.err
   python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_${wl}_${name}${mode}.json").read().strip().splitlines()[-1])
    print("$wl $name '$mode':", d["ms_per_step"])
except Exception as e:
    print("$wl $name '$mode' FAILED", e); print(open("gpurun_out/${tag}_${wl}_${name}${mode}.err").read()[-600:])
PY
  done
 done
done
