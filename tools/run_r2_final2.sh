# closing numbers after the last changes of the round (pool allocator, packed red-black, second team of the single-launch solver)
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref bench rc=$?"
timeout 1500 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --solves > gpurun_out/final_solves.json 2> gpurun_out/final_solves.err; echo "solves rc=$?"; cp gpurun_out/solves.md gpurun_out/final_solves.md
