#!/bin/bash
# A/B timing of library variants (csrc/Makefile `variant`) on the north-star bench's device-resident loop.
# usage: bash tools/gpu_variants2.sh <tag> <reads> <name> [<name> ...]   ("base" = the product library)
TAG=$1; READS=$2; shift; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset S5B_LIBRARY; else export S5B_LIBRARY=$PWD/slow5tools_b200/libslow5b200_$v.so; fi
  timeout 300 python bench.py --reads $READS --steps 5 --profile > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$v.json").read().splitlines()[-1]); s=d["stage_ms"]
    print("$v", "enc %.2f dec %.2f | inflate %.2f deflate %.2f sigdec %.2f sigenc %.2f pack %.2f image %.2f"%(d["encode_ms"],d["decode_ms"],s["record_depress"],s["record_press"],s["signal_depress"],s["signal_press"],s["pack"],s["image"]))
except Exception as e: print("$v","failed",e, open("gpurun_out/${TAG}_$v.err").read()[-600:])
PY
done
