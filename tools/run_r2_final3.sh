#!/bin/bash
# r2 re-entry closing run on one GPU: whole GPU suite, smoke, both bench arms, launch list, ncu rows of the blocks ordering
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/tests_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref bench rc=$?"
timeout 1200 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/final_bench_n1.json
python -c "
import json
d=json.load(open('gpurun_out/final_bench_n1.json'))
print(json.dumps(d['e2e']['solves'], indent=0))
print({k:(v.get('iters'), round(v.get('ms_per_iter',0),3), round(v.get('precapply_ms_per_apply',0),4)) for k,v in d['solve'].items() if isinstance(v,dict) and 'convdiff256' in k})
"
timeout 300 python tools/sgs_blocks_probe.py 64 128 256 384 > gpurun_out/blk_probe_final.log 2>&1; grep -v "staging" gpurun_out/blk_probe_final.log | head -80
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,smsp__warps_active.avg.per_cycle_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 200 python tools/sgs_one_blocks.py 256 > gpurun_out/sgs_one_blocks_256.log 2>&1 && timeout 600 ncu --metrics $M --clock-control none -k regex:sgs_warpflow -s 6 -c 2 --csv --log-file gpurun_out/ncu_sgs_blocks_256.csv python tools/sgs_one_blocks.py 256 > gpurun_out/ncu_sgs_blocks_256.log 2>&1
cat gpurun_out/sgs_one_blocks_256.log; tail -3 gpurun_out/ncu_sgs_blocks_256.csv | cut -c1-400
