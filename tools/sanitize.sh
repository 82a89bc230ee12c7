#!/bin/bash
# compute-sanitizer passes over the hot path (run under gpurun): memcheck and racecheck of the tile kernel's mbarrier
# record pipeline / warp-private mirrored stores and of the in-place panel / triangular-solve steps.
# Output: gpurun_out/<tag>_sanitizer_{memcheck,racecheck}.log (summaries are copied into profiles/ by hand).
TAG=${1:-r02}
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_case.py 10 \
      > gpurun_out/${TAG}_sanitizer_${TOOL}.log 2>&1
  echo "$TOOL rc=$?" >> gpurun_out/${TAG}_sanitizer_${TOOL}.log
  tail -4 gpurun_out/${TAG}_sanitizer_${TOOL}.log
done
