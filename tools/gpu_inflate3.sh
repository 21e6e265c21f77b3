#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_inflate_gpu.py tests/test_recode_gpu.py tests/test_configs_gpu.py -x -q > gpurun_out/i3_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/i3_pytest.log
tail -5 gpurun_out/i3_pytest.log
timeout 600 python bench.py --reads 1000000 --steps 5 --warmup 3 --profile 2>/dev/null | tail -1 > gpurun_out/i3_profile.json; cat gpurun_out/i3_profile.json
