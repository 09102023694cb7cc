#!/bin/bash
# bench.py at N GPUs with and without the overlapped halo exchange; prints the weak SpMV and strong CG numbers
N=${1:-8}
for ov in 1 0; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$ov \
      bench.py --gpus $N --steps 30 --warmup 3 --halo-overlap $ov 2>/dev/null > gpurun_out/bench${N}_ov$ov.json
  python - "$ov" "$N" <<'PY'
import json, sys
ov, N = sys.argv[1], sys.argv[2]
d = json.loads([l for l in open(f"gpurun_out/bench{N}_ov{ov}.json") if l.startswith("{")][-1])
s = d["solve"]
print("N", N, "overlap", ov, "| weak spmv ms/step", round(d["ms_per_step"], 4), "value", round(d["value"]), "| cg weak ms/iter", round(s["cg_jacobi_poisson"]["ms_per_iter"], 3),
      "| strong 1024^3 ms/iter", round(s["cg_jacobi_poisson_strong"]["ms_per_iter"], 3), s["cg_jacobi_poisson_strong"]["grid_per_gpu"], flush=True)
PY
done
