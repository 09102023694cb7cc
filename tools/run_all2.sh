mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu -rx > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests.log
tail -15 gpurun_out/gpu_tests.log
S3D_WRITE_HISTORY=gpurun_out/cg1024_history.json timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'frac',d['roofline']['frac'], 'e2e',d['e2e']['value'], d['e2e']['ms_per_step'])
print(d['solve'].get('error')); print(d['e2e'].get('solves'))
"
ls -la gpurun_out/cg1024_history.json
