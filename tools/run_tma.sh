#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "spmv" > gpurun_out/tma_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tma_tests.log
timeout 200 python tools/sweep.py --sizes 256,512 --reps 20 > gpurun_out/tma_sweep.log 2>&1; echo "sweep rc=$?"; grep -i "spmv" gpurun_out/tma_sweep.log | head -20
