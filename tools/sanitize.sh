#!/bin/bash
# compute-sanitizer memcheck + racecheck over the hot kernels on small windows; logs to gpurun_out/
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "== $tool rc=$? ==" >> gpurun_out/sanitizer_$tool.txt
  tail -4 gpurun_out/sanitizer_$tool.txt
done
