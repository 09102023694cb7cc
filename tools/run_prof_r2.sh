mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 300 python tools/profile_r2.py 512 > gpurun_out/profile_r2.log 2>&1 || { tail -5 gpurun_out/profile_r2.log; exit 1; }
timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_ncu_kernels_512.csv python tools/profile_r2.py 512 > gpurun_out/ncu_r2a.log 2>&1; echo "ncu kernels rc=$?"
timeout 600 python bench.py --steps 2 --warmup 3 --no-solve --no-cpu > gpurun_out/b.json 2> gpurun_out/b.err || { echo bench failed; tail -5 gpurun_out/b.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-solve --no-cpu > gpurun_out/ncu_r2b.log 2>&1; echo "ncu launches rc=$?"
timeout 100 python tools/sgs_one4.py 128 > gpurun_out/sgs_one4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:sgs_stream -s 4 -c 2 -o gpurun_out/r2_sgs_stream_v6_128 python tools/sgs_one4.py 128 > gpurun_out/ncu_r2c.log 2>&1; echo "ncu stream rc=$?"
wc -l gpurun_out/r2_ncu_kernels_512.csv gpurun_out/r2_bench_launches.csv; ls -la gpurun_out/*.ncu-rep
