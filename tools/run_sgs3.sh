mkdir -p gpurun_out
timeout 120 python tools/sgs_debug.py > gpurun_out/sgs_debug.log 2>&1; echo "rc=$?" >> gpurun_out/sgs_debug.log; grep -c "equal True" gpurun_out/sgs_debug.log; grep "equal False" gpurun_out/sgs_debug.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "sgs" > gpurun_out/sgs_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/sgs_tests.log
tail -5 gpurun_out/sgs_tests.log
timeout 100 python tools/sgs_one4.py 128 > gpurun_out/sgs_one4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:sgs_stream -s 4 -c 2 -o gpurun_out/r2_sgs_stream_128 python tools/sgs_one4.py 128 > gpurun_out/ncu_one4.log 2>&1
cat gpurun_out/sgs_one4.log; tail -3 gpurun_out/ncu_one4.log
