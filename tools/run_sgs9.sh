mkdir -p gpurun_out
SGS_PROF=1 timeout 300 python tools/sgs_probe.py 256 512 > gpurun_out/sgs_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/sgs_probe.log
cat gpurun_out/sgs_probe.log
