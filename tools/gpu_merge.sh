#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_merge_split.py tests/test_recode_gpu.py tests/test_view_gpu.py -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/m_pytest.log
tail -15 gpurun_out/m_pytest.log
