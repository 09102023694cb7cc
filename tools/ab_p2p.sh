#!/bin/bash
# Correctness (tests/mgpu_check.py, default = peer-memory halo) and A/B of the halo exchange on N GPUs: peer-memory windows vs NCCL send/recv.
N=${1:-2}
mkdir -p gpurun_out
for p2p in 1 0 1 0; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$p2p \
      bench.py --gpus $N --steps 30 --warmup 3 --strong 1024 --strong-iters 40 --cg-max-it ${CGMAX:-20000} --halo-p2p $p2p 2>gpurun_out/ab_p2p_$p2p.err > /tmp/ab_$p2p.json || tail -5 gpurun_out/ab_p2p_$p2p.err
  python - "$p2p" <<'PY'
import json, sys
v = sys.argv[1]
d = json.loads([l for l in open(f"/tmp/ab_{v}.json") if l.startswith("{")][-1])
cg = d["solve"]["cg_jacobi_poisson"]; st = d["solve"].get("cg_jacobi_poisson_strong", {})
print("halo_p2p", v, "| spmv step ms", round(d["ms_per_step"], 4), "| value GB/s", round(d["value"]), "| cg 512^3/GPU ms/iter", round(cg["ms_per_iter"], 3), "iters", cg["iters"],
      "| strong 1024^3 ms/iter", round(st.get("ms_per_iter", 0), 3), flush=True)
PY
done
