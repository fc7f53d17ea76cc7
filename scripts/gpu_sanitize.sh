#!/bin/bash
# compute-sanitizer over scripts/sanitize_scenes.py (every device path on small scenes), three tools.
tag=${1:-r2f}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool python scripts/sanitize_scenes.py > gpurun_out/sanitizer_${tag}_$tool.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/sanitizer_${tag}_$tool.log
done
