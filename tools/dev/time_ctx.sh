#!/bin/bash
# dev: what the CLI's start-up is made of
CLI=slow5tools_b200/bin/slow5tools-b200
F=tests/golden/fixtures/exp_1_lossless_zlib_svb_v0.2.0.blow5
nvidia-smi --query-gpu=persistence_mode --format=csv,noheader | head -2
for i in 1 2 3; do ( time S5B_TIMING=1 $CLI view $F -c none -s none -o /dev/shm/ctx_out.blow5 ) 2>&1 | grep -v "^$" | grep -v "^user\|^sys"; done
echo "--- CUDA_MODULE_LOADING=EAGER"
( time CUDA_MODULE_LOADING=EAGER S5B_TIMING=1 $CLI view $F -c none -s none -o /dev/shm/ctx_out.blow5 ) 2>&1 | grep -v "^$" | grep -v "^user\|^sys"
rm -f /dev/shm/ctx_out.blow5
