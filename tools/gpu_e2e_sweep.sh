#!/bin/bash
# dev: e2e sensitivity to the pipeline chunk size
mkdir -p gpurun_out
for mb in 8 16 32 64 128 256; do
  S5B_CHUNK_MB=$mb timeout 300 python bench.py --no-zlib --steps 20 > gpurun_out/e2e_$mb.json 2> gpurun_out/e2e_$mb.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/e2e_$mb.json")); print("chunk $mb MB: e2e %.4g reads/s  value %.4g"%(d["e2e"]["value"], d["value"]))
except Exception as e: print("$mb failed", e)
PY
done
