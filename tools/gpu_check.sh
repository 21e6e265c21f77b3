#!/bin/bash
# One GPU-box visit: parity tests, bench line, reference arm, ncu launch list and one --set full capture per kernel.
# usage (under gpurun): bash tools/gpu_check.sh <tag> [skip-tests]
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
if [ "$2" != "skip-tests" ]; then
  timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -3 gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench exit $?"; tail -c 600 gpurun_out/${TAG}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
KR='regex:svbzd_|inflate|deflate|zstd|exzd'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 200 --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --profile > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "$KR" -s 8 -c 12 \
   -o gpurun_out/${TAG}_full -f python bench.py --steps 2 --warmup 3 --profile > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv -k regex:svbzd_ 2>/dev/null | gzip > gpurun_out/${TAG}_svbzd_source.csv.gz
# gpurun brings back at most 64 MiB: the report itself stays on the box unless it is small
[ $(stat -c %s gpurun_out/${TAG}_full.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/${TAG}_full.ncu-rep
ls -la gpurun_out | tail -20
