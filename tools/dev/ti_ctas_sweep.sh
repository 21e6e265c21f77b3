#!/bin/bash
# inflate_thread_kernel: CTAs (of 2 warps) per SM against the 32-record rounds of a single 1 M-record chunk
# (libraries built with `make variant NAME=ti<c> VSRC=inflate_thread_kernels DEFS=-DS5B_TI_CTAS=<c>`)
echo "base (10)"; python bench.py --profile --steps 3 --warmup 3 2>/dev/null | tail -1
for c in 9 11 12; do
  echo "ti$c"; S5B_LIBRARY=$PWD/slow5tools_b200/libslow5b200_ti$c.so python bench.py --profile --steps 3 --warmup 3 2>/dev/null | tail -1
done
