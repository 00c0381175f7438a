#!/bin/bash
python scripts/latency_probe.py 2>&1 | tail -7 | tee gpurun_out/r02am_latency.log
echo "--- copy path"; PRGPU_NO_HOST_WINDOW=1 python scripts/latency_probe.py 2>&1 | tail -7 | tee -a gpurun_out/r02am_latency.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02am_launches.csv python scripts/latency_probe.py > /dev/null 2>&1
