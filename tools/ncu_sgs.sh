#!/bin/bash
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__warps_active.avg.per_cycle_active,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum
for P in 1 2; do
timeout 200 python tools/sgs_one.py 512 $P > gpurun_out/sgs_one_p$P.log 2>&1 || exit 1
timeout 600 ncu --metrics $M --clock-control none -k regex:sgs_warpflow -s 6 -c 2 --csv --log-file gpurun_out/ncu_sgs_p$P.csv python tools/sgs_one.py 512 $P > gpurun_out/ncu_sgs_p$P.log 2>&1
done
cat gpurun_out/sgs_one_p*.log
