mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_gpu.py -x -q -m gpu > gpurun_out/host_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/host_tests.log
tail -5 gpurun_out/host_tests.log
timeout 900 python bench.py --solves --solves-large 0 > gpurun_out/solves.json 2> gpurun_out/solves.err; echo "solves rc=$?"
cat gpurun_out/solves.md
