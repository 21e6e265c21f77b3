#!/bin/bash
TAG=${1:-exzd}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_exzd_gpu.py tests/test_view_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -30 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 50 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench.json"))
    print(json.dumps(d["exzd_stage"], indent=1))
    print("svb enc/dec ms", d["encode_ms"], d["decode_ms"])
except Exception as e: print("bench failed", e, open("gpurun_out/${TAG}_bench.err").read()[-1500:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:exzd_ -c 2 \
   -o gpurun_out/${TAG}_full -f python bench.py --steps 2 --warmup 3 --profile > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_source.csv.gz
[ $(stat -c %s gpurun_out/${TAG}_full.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/${TAG}_full.ncu-rep
ls -la gpurun_out | tail -8
