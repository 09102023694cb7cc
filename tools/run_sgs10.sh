mkdir -p gpurun_out
S3D_SGS_LINELOG=1 timeout 300 python tools/sgs_lines.py 2048,64,8 > gpurun_out/sgs_lines.log 2>&1
cat gpurun_out/sgs_lines.log | head -40
