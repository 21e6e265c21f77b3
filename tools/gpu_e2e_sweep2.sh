#!/bin/bash
# e2e (host-buffer) throughput of the north-star step against the pipeline's knobs: lanes x chunk records
# usage: bash tools/gpu_e2e_sweep2.sh "<lanes list>" "<chunk list>"
mkdir -p gpurun_out
for lanes in ${1:-3}; do for chunk in ${2:-8192 16384 32768 65536}; do
  S5B_RECODE_LANES=$lanes S5B_RECODE_CHUNK=$chunk timeout 300 python bench.py --e2e-only --steps 2 --warmup 3 > gpurun_out/e2e_l${lanes}_c${chunk}.json 2> gpurun_out/e2e_l${lanes}_c${chunk}.err
  echo "lanes $lanes chunk $chunk: $(tail -1 gpurun_out/e2e_l${lanes}_c${chunk}.json)"
done; done
