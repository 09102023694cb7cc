#!/bin/bash
# A/B of the overlapped halo exchange on N GPUs: bench.py with --halo-overlap 1 / 0, alternating.
N=${1:-2}
for ov in 1 0 1 0; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$ov \
      bench.py --gpus $N --steps 30 --warmup 3 --strong 0 --halo-overlap $ov 2>/dev/null > /tmp/ab_$ov.json
  python - "$ov" <<'PY'
import json, sys
ov = sys.argv[1]
d = json.loads([l for l in open(f"/tmp/ab_{ov}.json") if l.startswith("{")][-1])
cg = d["solve"]["cg_jacobi_poisson"]
print("overlap", ov, "spmv ms/step", round(d["ms_per_step"], 4), "cg ms/iter", round(cg["ms_per_iter"], 3), "iters", cg["iters"], flush=True)
PY
done
