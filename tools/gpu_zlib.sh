#!/bin/bash
# entropy-kernel iteration: parity tests + the bench's zlib/zstd stage timings
TAG=${1:-zl}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_inflate_gpu.py tests/test_deflate_gpu.py tests/test_zstd_gpu.py tests/test_zstd_encode_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read().splitlines()[-1]); z=d["zlib_stage"]
    print("deflate %.2f inflate %.2f zstd enc %.2f dec %.2f ms"%(z["deflate_ms"],z["inflate_ms"],z["zstd"]["encode_ms"],z["zstd"]["decode_ms"]))
except Exception as e: print("bench failed", e, open("gpurun_out/${TAG}_bench.err").read()[-1500:])
PY
