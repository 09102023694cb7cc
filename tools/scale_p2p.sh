#!/bin/bash
# N-GPU check: tests/mgpu_check.py (correctness over N slabs, peer-memory halo), then bench.py with the peer-memory halo and a
# short NCCL run beside it
N=${1:-8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/mgpu_check.py > gpurun_out/mgpu_check_n$N.log 2>&1; echo "mgpu_check rc=$?"
grep "PASS\|FAIL" gpurun_out/mgpu_check_n$N.log | tail -4
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus $N --steps 30 --warmup 3 --halo-p2p 1 2>gpurun_out/bench_n${N}_p2p1.err > gpurun_out/bench_n${N}_p2p1.json; echo "bench p2p rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 30 --warmup 3 --halo-p2p 0 --no-solve 2>gpurun_out/bench_n${N}_p2p0.err > gpurun_out/bench_n${N}_p2p0.json; echo "bench nccl rc=$?"
python - "$N" <<'PY'
import json, sys
N = sys.argv[1]
for v in ("1", "0"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_n{N}_p2p{v}.json") if l.startswith("{")][-1])
    except Exception as e:
        print("p2p", v, "no line:", e); continue
    s = d.get("solve", {})
    print("N", N, "halo_p2p", v, "| weak spmv ms/step", round(d["ms_per_step"], 4), "value", round(d["value"]), "frac of measured", round(d["value_frac_of_measured_peak"], 3),
          "| cg weak ms/iter", round(s.get("cg_jacobi_poisson", {}).get("ms_per_iter", 0), 3), "iters", s.get("cg_jacobi_poisson", {}).get("iters"),
          "| strong 1024^3 ms/iter", round(s.get("cg_jacobi_poisson_strong", {}).get("ms_per_iter", 0), 3), flush=True)
PY
