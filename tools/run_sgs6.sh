mkdir -p gpurun_out
timeout 300 python tools/sgs_lines.py 512,4,8 2048,4,8 2048,8,8 2048,4,16 2048,16,8 2048,4,32 2048,16,32 > gpurun_out/sgs_lines.log 2>&1; echo "rc=$?" >> gpurun_out/sgs_lines.log
cat gpurun_out/sgs_lines.log
