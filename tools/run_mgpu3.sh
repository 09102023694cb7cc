N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "overlap or blocks" > gpurun_out/ov_tests_n$N.log 2>&1; echo "one-device overlap/blocks tests rc=$?"; tail -3 gpurun_out/ov_tests_n$N.log
bash tools/run_mgpu2.sh $N
