set -e
cd /root/repo
python - <<'PY'
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tools')
import bench_view, numpy as np
from slow5tools_b200 import synth
sig = synth.nanopore_signal(200000*4096, seed=42).numpy()
bench_view.write_blow5('/dev/shm/raw.blow5', sig, 200000, 4096)
PY
B=slow5tools_b200/bin/slow5tools-b200
for i in 1 2; do time env S5B_TIMING=1 $B view /dev/shm/raw.blow5 -o /dev/shm/z.blow5; done
for i in 1 2; do time env S5B_TIMING=1 $B view /dev/shm/z.blow5 -c none -s none -o /dev/shm/back.blow5; done
cmp /dev/shm/back.blow5 /dev/shm/raw.blow5 && echo roundtrip ok
time cat /dev/shm/raw.blow5 > /dev/null
rm -f /dev/shm/raw.blow5 /dev/shm/z.blow5 /dev/shm/back.blow5
