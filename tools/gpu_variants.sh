#!/bin/bash
# dev: A/B timing of kernel-variant libraries (csrc/Makefile `variant`) on the GPU box.
# usage: bash tools/gpu_variants.sh <tag> <name> [<name> ...]   ("base" = the product library)
TAG=$1; shift
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_svbzd_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
for v in "$@"; do
  if [ "$v" = base ]; then unset S5B_LIBRARY; else export S5B_LIBRARY=$PWD/slow5tools_b200/libslow5b200_$v.so; fi
  if [ "$v" != base ]; then timeout 300 python -m pytest tests/test_svbzd_gpu.py -x -q 2>&1 | tail -1; fi
  timeout 300 python bench.py --no-zlib --steps 100 > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_$v.json"))
    print("$v", "enc %.4f dec %.4f ms  frac %.3f/%.3f  e2e %.3g"%(d["encode_ms"],d["decode_ms"],d["roofline"]["encode_frac"],d["roofline"]["decode_frac"],d["e2e"]["value"]))
except Exception as e: print("$v","failed",e, open("gpurun_out/${TAG}_bench_$v.err").read()[-600:])
PY
done
