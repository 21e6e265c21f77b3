#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deflate_gpu.py tests/test_recode_gpu.py -x -q > gpurun_out/d3_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/d3_pytest.log
tail -5 gpurun_out/d3_pytest.log
timeout 600 python bench.py --reads 1000000 --steps 5 --warmup 3 --profile 2>/dev/null | tail -1 > gpurun_out/d3_profile.json; cat gpurun_out/d3_profile.json
