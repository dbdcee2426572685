#!/bin/bash
# compute-sanitizer passes over the clustering / build GPU tests (SURVEY section 5 aux: race detection).
# Usage (on the GPU box): bash scripts/sanitize.sh [tag]   -> gpurun_out/sanitize_<tag>_{memcheck,racecheck}.txt
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
SEL=${SEL:-"cluster or build or scan"}
for tool in memcheck racecheck; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests -m gpu -x -q -k "$SEL" -p no:cacheprovider > $out/sanitize_${tag}_${tool}.txt 2>&1
  echo "exit $?" >> $out/sanitize_${tag}_${tool}.txt
  tail -5 $out/sanitize_${tag}_${tool}.txt
done
