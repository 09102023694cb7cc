# Round-2 closing measurements on one B200: GPU tests, bench (both arms), kernel metrics, launch list, one full capture per kernel of interest
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/final_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/final_gpu_tests.log
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref bench rc=$?"
timeout 1500 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --solves > gpurun_out/final_solves.json 2> gpurun_out/final_solves.err; echo "solves rc=$?"; cp gpurun_out/solves.md gpurun_out/final_solves.md
timeout 900 python bench.py --sweep > gpurun_out/final_sweep.json 2> gpurun_out/final_sweep.err; echo "sweep rc=$?"; cp gpurun_out/bw_sweep.md gpurun_out/final_bw_sweep.md 2>/dev/null
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 300 python tools/profile_r2.py 512 > gpurun_out/profile_r2.log 2>&1 || { tail -5 gpurun_out/profile_r2.log; exit 1; }
timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_ncu_kernels_512.csv python tools/profile_r2.py 512 > gpurun_out/ncu_r2a.log 2>&1; echo "ncu kernels rc=$?"
timeout 600 python bench.py --steps 2 --warmup 3 --no-solve --no-cpu > gpurun_out/b.json 2> gpurun_out/b.err || { echo bench failed; tail -5 gpurun_out/b.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-solve --no-cpu > gpurun_out/ncu_r2b.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_march -s 6 -c 1 -o gpurun_out/r2_spmv_march_512 -f python bench.py --steps 2 --warmup 3 --no-solve --no-cpu > gpurun_out/ncu_r2c.log 2>&1; echo "ncu spmv full rc=$?"
timeout 100 python tools/sgs_tile_mail_probe.py 256 > gpurun_out/tile_mail_256.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:sgs_warpflow -s 8 -c 2 -o gpurun_out/r2_sgs_tile_mail_256 -f python tools/sgs_tile_mail_probe.py 256 > gpurun_out/ncu_r2d.log 2>&1; echo "ncu sgs full rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
