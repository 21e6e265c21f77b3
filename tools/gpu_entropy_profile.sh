#!/bin/bash
# ncu --set full + source page of the zlib kernels (deflate_*_kernel, inflate_kernel) on the north-star bench's batch.
# usage (under gpurun): bash tools/gpu_entropy_profile.sh <tag> [reads] [kernel regex]
TAG=${1:-ent}; READS=${2:-100000}; KR=${3:-deflate_count|deflate_tree|deflate_emit|inflate_kernel}
mkdir -p gpurun_out
# one launch of every kernel, taken after the warm-up steps (-s skips the first matches)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KR" -s ${SKIP:-8} -c ${COUNT:-4} \
   -o gpurun_out/${TAG}_full -f python bench.py --steps 2 --warmup 3 --profile --reads $READS > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_source.csv.gz
[ $(stat -c %s gpurun_out/${TAG}_full.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/${TAG}_full.ncu-rep
tail -5 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out | grep ${TAG}
