mkdir -p gpurun_out
timeout 100 python tools/sgs_one4.py 512 > gpurun_out/sgs_one4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:sgs_stream -s 4 -c 1 -o gpurun_out/r2_sgs_stream_512 python tools/sgs_one4.py 512 > gpurun_out/ncu_one4.log 2>&1
cat gpurun_out/sgs_one4.log; tail -3 gpurun_out/ncu_one4.log
