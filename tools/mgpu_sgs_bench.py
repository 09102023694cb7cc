#!/usr/bin/env python
"""BiCGSTAB + SGS on a z-slab decomposition (torchrun, one rank per GPU): exact lexicographic sweep through the slabs vs red-black.
Grid: n x n x (n * world) convection-diffusion, manufactured right-hand side."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from struct3d_b200 import capi, hostapi as H
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
obj = [capi.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(obj, src=0)
H.init(local, rank, world, obj[0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m, _ = H._capture_stdout(lambda: H.Mesh((n + 2, n + 2, n * world + 2)))
P, _ = H._capture_stdout(lambda: H.Pde("convdiff", (0.577350269189626, 0.7, 0.3), 0.1))
A, b, u = P.lhs(m), P.vector(m, "manufactured_rhs"), H.Vec(m)
for order in (("lexicographic", "redblack") if n >= 512 else ("lexicographic", "redblack", "lexicographic", "redblack")):
    H.set_options({"-s3d_ksp_type": "bcgs", "-s3d_pc_type": "sgs", "-ksp_rtol": 1e-6, "-ksp_max_it": 5000, "-s3d_pc_apply_sweeps": 1,
                   "-s3d_thread_chunk_size": 1024, "-s3d_sgs_ordering": order})
    (s, _) = H._capture_stdout(lambda: H.Solver(A)); s.update(); H.sync(); dist.barrier()
    H.vecset(0.0, u)
    t0 = time.perf_counter(); info, _ = H._capture_stdout(lambda: s.apply(b, u)); H.sync(); dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{world} GPUs, {n}x{n}x{n * world}, bcgs + SGS {order}: {info['iters']} iterations, {t.item():.4f} s, {t.item() * 1e3 / info['iters']:.3f} ms/iteration, "
              f"preconditioner {info['precapplywtime']:.4f} s, converged {info['converged']}", flush=True)
    del s
dist.barrier(); dist.destroy_process_group()
