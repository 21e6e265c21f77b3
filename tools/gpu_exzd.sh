#!/bin/bash
TAG=${1:-exzd}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_exzd_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -30 gpurun_out/${TAG}_pytest.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_exzd_gpu.py -x -q -k "golden or ragged or malformed" > gpurun_out/${TAG}_sanitizer.log 2>&1
tail -5 gpurun_out/${TAG}_sanitizer.log
