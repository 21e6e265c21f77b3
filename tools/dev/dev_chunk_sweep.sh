#!/bin/bash
# device-resident transcoder: records per chunk against the 32-record rounds of the inflate kernel (2960 warps per B200)
for c in 262144 284160 378880 473600 1000000; do
  echo "chunk $c"; S5B_RECODE_DEV_CHUNK=$c python bench.py --profile --steps 3 --warmup 3 2>/dev/null | tail -1
done
for mb in 4096 9000; do
  echo "chunk 1000000 mb $mb"; S5B_RECODE_DEV_CHUNK=1000000 S5B_RECODE_DEV_CHUNK_MB=$mb python bench.py --profile --steps 3 --warmup 3 2>&1 | tail -1
done
