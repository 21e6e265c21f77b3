#!/bin/bash
# deflate / inflate iteration: codec parity tests, record-path tests, a short north-star bench
TAG=${1:-df}; READS=${2:-200000}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_deflate_gpu.py tests/test_inflate_gpu.py tests/test_recode_gpu.py tests/test_configs_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -12 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --reads $READS --steps 5 --profile > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1200 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
