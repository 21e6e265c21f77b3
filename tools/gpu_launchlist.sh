#!/bin/bash
# launch list of the north-star step: every kernel of the library, gpu__time_duration only.  usage: bash tools/gpu_launchlist.sh <tag> [reads]
TAG=${1:-ll}; READS=${2:-1000000}
mkdir -p gpurun_out
KR='regex:s5b|svbzd|inflate|deflate|rec_|image_|zstd|scan_|exzd|recode_|rebase|ascii|gather_copy|sig_extract|qts'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 2500 --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --profile --reads $READS > gpurun_out/${TAG}_launches.log 2>&1
tail -1 gpurun_out/${TAG}_launches.log | cut -c1-300
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/${TAG}_launches.csv") if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    t[r[ki]][0] += 1; t[r[ki]][1] += v
tot = sum(v[1] for v in t.values())
with open("gpurun_out/${TAG}_launch_summary.txt", "w") as f:
    for k, v in sorted(t.items(), key=lambda kv: -kv[1][1]):
        line = "%-60s launches %5d  total %9.3f ms  per launch %8.4f ms  share %5.1f %%" % (k[:60], v[0], v[1], v[1] / v[0], 100 * v[1] / tot)
        print(line); f.write(line + "\n")
PY
