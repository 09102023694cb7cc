# A/B at N GPUs: weak SpMV + e2e (no solves) with the side-stream push and the peer all-reduce on / off
N=${1:-2}
mkdir -p gpurun_out
for t in "halo_push_side=1" "halo_push_side=0"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-solve --no-cpu --tune $t > gpurun_out/ab_$t.json 2> gpurun_out/ab_$t.err
  python -c "
import json
d=json.load(open('gpurun_out/ab_$t.json'))
print('$t', 'spmv ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['value'])
"
done
