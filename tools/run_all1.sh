mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests.log
tail -6 gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('cpu_baseline'))
print(json.dumps(d['solve'], indent=1)[:3000])
print(d['e2e'].get('solves'))
"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-900
nproc; lscpu | grep -E "Model name|avx512f" | cut -c1-200 | head -3
