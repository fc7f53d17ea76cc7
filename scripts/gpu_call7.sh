#!/bin/bash
# Round-2 session G (2 GPUs): ViewShardedBackward over NCCL vs the sequential reference, then the N = 1 / 2 bench lines.
tag=${1:-r2n}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider --timeout 600 > gpurun_out/${tag}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${tag}_pytest_multi.log
python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err
show() { python - <<PY
import json
try:
    d=[l for l in open("gpurun_out/${tag}_$1.json").read().strip().splitlines() if l.startswith("{")][-1]
    d=json.loads(d)
    print("$1:", d["ms_per_step"], d["value"], d["config"]["timed_region"], "eager", d["config"]["eager_ms_per_step"], "e2e", (d.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/${tag}_$1.err").read()[-1500:])
PY
}
show n1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_n2.json 2> gpurun_out/${tag}_n2.err
show n2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-overlap > gpurun_out/${tag}_n2_noverlap.json 2> gpurun_out/${tag}_n2_noverlap.err
show n2_noverlap
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/${tag}_n2_ref.json 2> gpurun_out/${tag}_n2_ref.err
tail -c 600 gpurun_out/${tag}_n2_ref.json
