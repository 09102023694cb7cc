#!/bin/bash
# ncu --set full of the element-wise kernel (axpy, copy) in its segment order, 512^3
mkdir -p gpurun_out
timeout 200 python tools/ew_probe.py 512 > gpurun_out/ew_probe.log 2>&1 || exit 1
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --clock-control none --kernel-name-base demangled -k "regex:ew_chunk_kernel.*(OpAxpy|OpCopy|OpCgUpdate)" -c 72 -o /tmp/ew_full python tools/ew_probe.py 512 > gpurun_out/ncu_ew.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/ew_full.ncu-rep --page raw --csv > gpurun_out/ncu_ew_raw.csv 2>/dev/null; wc -c gpurun_out/ncu_ew_raw.csv
