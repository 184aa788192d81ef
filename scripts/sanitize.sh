#!/bin/bash
# compute-sanitizer memcheck + racecheck of the smoke run and of one small config-2 step (team barriers,
# __syncwarp-only stages: DESIGN.md 2); logs -> gpurun_out/<tag>_sanitizer_*.log
TAG=${1:-r2}
SMOKE='python -c "import __graft_entry__ as g; g.smoke()"'
STEP="python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline --no-extras --parity-signals 2"
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 bash -c "$SMOKE" > gpurun_out/${TAG}_sanitizer_${tool}_smoke.log 2>&1
  echo "$tool smoke rc=$?"; tail -3 gpurun_out/${TAG}_sanitizer_${tool}_smoke.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 $STEP > gpurun_out/${TAG}_sanitizer_${tool}_step.log 2>&1
  echo "$tool step rc=$?"; tail -3 gpurun_out/${TAG}_sanitizer_${tool}_step.log
done
