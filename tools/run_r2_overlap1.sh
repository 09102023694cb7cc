#!/bin/bash
# r2 re-entry: the overlapping-slab SGS on one device, then the whole GPU suite, smoke and the N = 1 bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k overlap > gpurun_out/ov_tests.log 2>&1; echo "overlap tests rc=$?"; tail -5 gpurun_out/ov_tests.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/tests_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench.json
