#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_check.py > gpurun_out/mgpu_check_n$N.log 2>&1; echo "mgpu_check rc=$?"
grep "PASS\|FAIL\|Error\|error\|Traceback" gpurun_out/mgpu_check_n$N.log | tail -30
