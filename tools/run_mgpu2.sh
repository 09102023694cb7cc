N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mgpu.py -x -q -m gpu -s > gpurun_out/mgpu_pytest_n$N.log 2>&1; echo "mgpu pytest rc=$?"
grep "PASS\|FAIL\|sgs-gate\|Error\|error\|Traceback\|passed\|failed" gpurun_out/mgpu_pytest_n$N.log | tail -40
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_n$N.err
python -c "
import json,sys
d=json.load(open('gpurun_out/bench_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','halo_check')}, 'e2e',d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('equals_device_resident_result'))
for k,v in d['solve'].items(): print(k, v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in ('seconds','iters','ms_per_iter','ms_per_iter_min','ms_per_iter_median','precapply_ms_per_apply','converged','grid','grid_total')})
"
