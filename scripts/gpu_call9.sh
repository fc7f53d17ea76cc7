#!/bin/bash
# Round-2 session I (2 GPUs): N = 2 bench, geometry half with and without the z pass.
tag=${1:-r2p}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_views.py -k "two_calls or graph or batch_of_views or error" -m gpu -q -x -p no:cacheprovider --timeout 600 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
show() { python - <<PY
import json
try:
    d=[l for l in open("gpurun_out/${tag}_$1.json").read().strip().splitlines() if l.startswith("{")][-1]
    d=json.loads(d)
    print("$1:", d["ms_per_step"], d["value"], d["config"]["timed_region"][:60], "eager", d["config"]["eager_ms_per_step"], "e2e", (d.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/${tag}_$1.err").read()[-1500:])
PY
}
DEODR_B200_GEOMETRY_Z=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_n2_z1.json 2> gpurun_out/${tag}_n2_z1.err
show n2_z1
DEODR_B200_GEOMETRY_Z=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_n2_z0.json 2> gpurun_out/${tag}_n2_z0.err
show n2_z0
DEODR_B200_GEOMETRY_Z=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_n2_z1b.json 2> gpurun_out/${tag}_n2_z1b.err
show n2_z1b
