#!/bin/bash
# stage times of one pipeline chunk in the flow of the step vs with an idle GPU around every stage
mkdir -p gpurun_out
for iso in 0 1; do
  if [ $iso = 1 ]; then export S5B_STAGE_ISOLATE=1; else unset S5B_STAGE_ISOLATE; fi
  python bench.py --steps 5 --warmup 3 --profile --reads 250000 2>/dev/null | tail -1 > gpurun_out/stage_probe_$iso.json
  echo "isolate=$iso"; cat gpurun_out/stage_probe_$iso.json
done
