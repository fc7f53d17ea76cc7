#!/bin/bash
tag=${1:-r2v}
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 100 compute-sanitizer --tool $tool python scripts/sanitize_scenes.py > gpurun_out/sanitizer_${tag}_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_${tag}_$tool.log
done
