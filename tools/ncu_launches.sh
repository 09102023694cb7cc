#!/bin/bash
# launch list of the default bench command (after it exited 0 without ncu)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/b.json 2> gpurun_out/b.err || { echo "bench failed"; tail -5 gpurun_out/b.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
