#!/bin/bash
mkdir -p gpurun_out
python tools/dev/svb_probe.py 250000 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:svbzd_decode -c 30 --csv --log-file gpurun_out/svb_probe_ncu.csv python tools/dev/svb_probe.py 250000 > /dev/null 2>&1
grep svbzd_decode gpurun_out/svb_probe_ncu.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -12 | tr '\n' ' '; echo
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:svbzd_decode -c 30 --csv --log-file gpurun_out/svb_probe_ncu2.csv python tools/dev/svb_probe.py 250000 > /dev/null 2>&1
echo "cache-control none:"; grep svbzd_decode gpurun_out/svb_probe_ncu2.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -12 | tr '\n' ' '; echo
