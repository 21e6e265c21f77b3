#!/bin/bash
# dev: timing breakdown of the view fast path (S5B_TIMING) on a synthetic file
R=${1:-300000}
python - <<PY
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import bench_view
from slow5tools_b200 import synth
bench_view.write_blow5('/dev/shm/tv_raw.blow5', synth.nanopore_signal($R*4096, seed=42).numpy(), $R, 4096)
PY
CLI=slow5tools_b200/bin/slow5tools-b200
for m in "zlib svb-zd" "zstd svb-zd"; do set -- $m
  echo "== encode -c $1 -s $2"; ( time S5B_TIMING=1 $CLI view -t 16 -K 20000 /dev/shm/tv_raw.blow5 -c $1 -s $2 -o /dev/shm/tv_z.blow5 ) 2>&1 | grep -v "^$" | tail -12
  echo "== decode"; ( time S5B_TIMING=1 $CLI view -t 16 -K 20000 /dev/shm/tv_z.blow5 -c none -s none -o /dev/shm/tv_back.blow5 ) 2>&1 | grep -v "^$" | tail -12
done
echo "== reference decode zstd"; ( time oracle/_ref/slow5tools_ref view -t 16 /dev/shm/tv_z.blow5 -c none -s none -o /dev/shm/tv_back.blow5 ) 2>&1 | tail -4
rm -f /dev/shm/tv_*.blow5
