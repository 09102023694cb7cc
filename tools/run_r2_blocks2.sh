#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "blocks or overlap or double_buffered or dataflow_variants or jacobi_and_sgs" > gpurun_out/blk_tests2.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/blk_tests2.log
S3D_PROBE_SHORT=1 timeout 600 python tools/sgs_blocks_probe.py 128 256 384 512 2>&1 | grep "staging areas" > gpurun_out/dbuf_probe.log; cat gpurun_out/dbuf_probe.log
