#!/bin/bash
# svb-zd kernel iteration on the GPU box: parity tests, A/B bench against the legacy kernels, ncu capture.
TAG=${1:-svb}
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_svbzd_gpu.py tests/test_view_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_svbzd_gpu.py -x -q -k "kat or ragged or adversarial" > gpurun_out/${TAG}_sanitizer.log 2>&1
tail -3 gpurun_out/${TAG}_sanitizer.log
S5B_SVBZD_LEGACY=1 timeout 300 python bench.py --no-zlib --steps 100 > gpurun_out/${TAG}_bench_legacy.json 2> gpurun_out/${TAG}_bench_legacy.err
timeout 300 python bench.py --no-zlib --steps 100 > gpurun_out/${TAG}_bench_new.json 2> gpurun_out/${TAG}_bench_new.err
python - <<PY
import json
for k in ("legacy","new"):
    try:
        d=json.load(open("gpurun_out/${TAG}_bench_%s.json"%k))
        print(k, "enc %.4f dec %.4f ms  frac %.3f/%.3f  e2e %.3g"%(d["encode_ms"],d["decode_ms"],d["roofline"]["encode_frac"],d["roofline"]["decode_frac"],d["e2e"]["value"]))
    except Exception as e: print(k,"failed",e, open("gpurun_out/${TAG}_bench_%s.err"%k).read()[-600:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svbzd_ -s 6 -c 2 \
   -o gpurun_out/${TAG}_full -f python bench.py --steps 2 --warmup 3 --profile --no-zlib > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_source.csv.gz
[ $(stat -c %s gpurun_out/${TAG}_full.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/${TAG}_full.ncu-rep
ls -la gpurun_out | tail
