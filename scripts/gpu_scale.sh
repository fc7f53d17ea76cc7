#!/bin/bash
# N-GPU bench of the headline workload, launched like the driver does (each arm under its own timeout).
n=${1:-2}; tag=${2:-scale}; shift; shift
mkdir -p gpurun_out
arms=("$@"); [ ${#arms[@]} -eq 0 ] && arms=("default:" "noverlap:--no-overlap")
for arm in "${arms[@]}"; do
  name=${arm%%:*}; flags=${arm#*:}
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 40 --warmup 5 --no-cpu-baseline $flags > gpurun_out/${tag}_n${n}_$name.json 2> gpurun_out/${tag}_n${n}_$name.err
  python - <<PY
import json
try:
    d=[l for l in open("gpurun_out/${tag}_n${n}_$name.json").read().strip().splitlines() if l.startswith("{")][-1]
    d=json.loads(d)
    print("N=$n $name:", d["ms_per_step"], d["value"], d["config"]["timed_region"], "eager", d["config"]["eager_ms_per_step"], "e2e", (d.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print("N=$n $name FAILED", e); print(open("gpurun_out/${tag}_n${n}_$name.err").read()[-1500:])
PY
done
