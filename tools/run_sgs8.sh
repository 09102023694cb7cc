mkdir -p gpurun_out
timeout 120 python tools/sgs_debug.py > gpurun_out/sgs_debug.log 2>&1; echo "rc=$?" >> gpurun_out/sgs_debug.log; grep -c "equal True" gpurun_out/sgs_debug.log; grep -A1 "equal False" gpurun_out/sgs_debug.log | head; tail -2 gpurun_out/sgs_debug.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "sgs" > gpurun_out/sgs_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/sgs_tests.log
tail -3 gpurun_out/sgs_tests.log
timeout 300 python tools/sgs_lines.py 2048,4,8 2048,8,8 2048,4,16 2048,16,32 > gpurun_out/sgs_lines.log 2>&1; echo "rc=$?" >> gpurun_out/sgs_lines.log
cat gpurun_out/sgs_lines.log
SGS_PROF=1 timeout 300 python tools/sgs_probe.py 64 128 256 512 > gpurun_out/sgs_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/sgs_probe.log
cat gpurun_out/sgs_probe.log
