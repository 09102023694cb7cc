#!/bin/bash
# SGS kernel check + timing on one GPU (bounded by timeouts: the dataflow kernels spin on flags / mailboxes)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "sgs" > gpurun_out/sgs_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/sgs_tests.log
tail -15 gpurun_out/sgs_tests.log
SGS_PROF=1 timeout 300 python tools/sgs_probe.py 64 128 256 512 > gpurun_out/sgs_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/sgs_probe.log
cat gpurun_out/sgs_probe.log
