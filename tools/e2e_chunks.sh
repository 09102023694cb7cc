# e2e (host vectors in / out) against the z-chunk count of the pipelined s3d_spmv_host
for c in 16 32 64 128 256; do
  timeout 200 python bench.py --steps 10 --warmup 3 --no-solve --no-cpu --e2e-chunks $c 2>/dev/null > gpurun_out/e2e_$c.json
  python -c "
import json
d=json.load(open('gpurun_out/e2e_$c.json'))
print('chunks $c: e2e %.3f ms/step  %.1f GB/s' % (d['e2e']['ms_per_step'], d['e2e']['value']))
"
done
