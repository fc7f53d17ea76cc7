#!/bin/bash
# The -m gpu suite + smoke + one bench line at HEAD.
tag=${1:-r2u}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 6 > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_c5.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], "eager", d["config"]["eager_ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["traffic"], d["roofline"]["frac"])
PY
