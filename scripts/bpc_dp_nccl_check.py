"""bPC data-parallel parity on real GPUs: torchrun --nproc-per-node 2 scripts/bpc_dp_nccl_check.py
Each rank runs jpc_b200.bpc_train_step on its contiguous row block (global-B normalisation, ONE NCCL all-reduce of both models'
gradients + energy pairs); rank 0 also runs the whole batch alone and compares gradients, energy pairs and its own rows of the
relaxed activities.  Clamped input and x = None, all three arithmetic modes; uneven shards on the fp32 path."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '.')
import jpc_b200 as J
import jpc_b200._train as T
from jpc_b200._train import shard_rows

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
for dtype, tol, B in (("f32", 1e-5, 257), ("bf16", 2e-2, 256), ("f32x3", 1e-5, 256)):
    for free in (False, True):
        J.set_compute_dtype(dtype)
        T._GLOBAL_ROWS.clear()
        d_x, w, d_y, depth = (128 if free else 64), 128, 32, 4
        td = J.make_mlp(0, d_x, w, depth, d_y, "tanh", use_bias=True)            # same seeds on every rank: replicated weights
        bu = J.make_mlp(1, d_y, w, depth, d_x, "tanh", use_bias=True)[::-1]
        g = torch.Generator().manual_seed(2)
        x = torch.randn(B, d_x, generator=g).cuda()
        y = torch.randn(B, d_y, generator=g).cuda()
        acts = [a + 0.3 * torch.randn(a.shape, generator=g).cuda() for a in J.init_activities_with_ffwd(td, x)]
        lo, hi = shard_rows(B, rank, world)
        kw = dict(n_steps=6, dt=0.05 * B, record_energies=True)
        A, gtd, gbu, E = J.bpc_train_step(td, bu, [a[lo:hi].contiguous() for a in acts], y[lo:hi].contiguous(),
                                          input=None if free else x[lo:hi].contiguous(), **kw)
        dist.barrier()
        if rank == 0:
            saved = T._world
            T._world = lambda: 1                                                  # the single-process run of the whole batch
            A1, rtd, rbu, E1 = J.bpc_train_step(td, bu, acts, y, input=None if free else x, **kw)
            T._world = saved
            rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
            ge = max(max(rel(p[1].weight, q[1].weight), rel(p[1].bias, q[1].bias)) for P, Q in ((gtd, rtd), (gbu, rbu)) for p, q in zip(P, Q))
            ae = max(rel(p, q[lo:hi]) for p, q in zip(A[:-1], A1[:-1]))
            ee = rel(E, E1)
            print(f"{dtype} x={'None' if free else 'given'} B={B} world={world}: gradient rel err {ge:.2e}, energy pairs {ee:.2e}, "
                  f"rank-0 activities {ae:.2e} (tol {tol:g})", flush=True)
            assert ge < tol and ee < max(tol, 1e-5) and ae < tol
dist.barrier()
if rank == 0:
    print("bPC data-parallel parity OK", flush=True)
os._exit(0)
