#!/bin/bash
# ncu --set full + source page of the zlib pair (deflate_kernel, inflate_kernel) on the bench's 100k x 4096 batch.
# usage (under gpurun): bash tools/gpu_entropy_profile.sh <tag>
TAG=${1:-ent}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^(s5b::)?(deflate_kernel|inflate_kernel)" -s 4 -c 2 \
   -o gpurun_out/${TAG}_full -f python bench.py --steps 2 --warmup 3 --profile --reads ${READS:-100000} > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_source.csv.gz
[ $(stat -c %s gpurun_out/${TAG}_full.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/${TAG}_full.ncu-rep
tail -5 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out | grep ${TAG}
