#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "blocks or overlap" > gpurun_out/blk_tests.log 2>&1; echo "blocks tests rc=$?"; tail -15 gpurun_out/blk_tests.log
timeout 600 python -m pytest tests/test_host_gpu.py -x -q -s -k "gate" > gpurun_out/blk_gate.log 2>&1; echo "gate tests rc=$?"; grep "sgs-gate\|passed\|failed\|Error" gpurun_out/blk_gate.log | tail -40
timeout 600 python tools/sgs_blocks_probe.py 64 128 256 384 512 > gpurun_out/blk_probe.log 2>&1; cat gpurun_out/blk_probe.log
