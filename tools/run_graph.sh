#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_gpu.py -x -q -m gpu > gpurun_out/host_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/host_tests.log
tail -5 gpurun_out/host_tests.log
timeout 600 python tools/graph_probe.py > gpurun_out/graph_probe.md 2> gpurun_out/graph_probe.err; echo "probe rc=$?"
grep "^|" gpurun_out/graph_probe.md; tail -3 gpurun_out/graph_probe.err
