#!/bin/bash
# round-2 GPU visit: record-path tests, (optionally) the whole GPU suite, the north-star bench and its reference arm.
# usage (under gpurun): bash tools/gpu_r2.sh <tag> [all|recode|none] [bench reads]
TAG=${1:-r2}; WHAT=${2:-recode}; READS=${3:-1000000}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
(nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA node\(s\)") >> gpurun_out/${TAG}_gpu.txt 2>&1
if [ "$WHAT" = all ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
elif [ "$WHAT" = recode ]; then
  timeout 600 python -m pytest tests/test_recode_gpu.py tests/test_view_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
fi
[ -f gpurun_out/${TAG}_pytest.log ] && tail -15 gpurun_out/${TAG}_pytest.log
if [ "$READS" != 0 ]; then
  timeout 900 python bench.py --reads $READS > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
  echo "bench exit $?"; tail -c 3000 gpurun_out/${TAG}_bench_n1.json; tail -5 gpurun_out/${TAG}_bench_n1.err
  timeout 400 python bench.py --impl reference --steps 5 --warmup 3 --reads $READS > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
  tail -c 600 gpurun_out/${TAG}_bench_ref.json
fi
