mkdir -p gpurun_out
timeout 120 python tools/sgs_debug.py > gpurun_out/sgs_debug.log 2>&1; echo "rc=$?" >> gpurun_out/sgs_debug.log; cat gpurun_out/sgs_debug.log
SGS_PROF=1 timeout 300 python tools/sgs_probe.py 64 256 512 > gpurun_out/sgs_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/sgs_probe.log
cat gpurun_out/sgs_probe.log
