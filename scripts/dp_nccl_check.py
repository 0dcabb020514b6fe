"""Data-parallel parity on real GPUs: torchrun --nproc-per-node 2 scripts/dp_nccl_check.py
Each rank runs make_pc_step on its contiguous row block of one batch (global-B normalisation, ONE NCCL all-reduce of the
gradient arena); rank 0 also runs the whole batch alone and compares updated weights, energies and loss."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '.')
import jpc_b200 as J
from jpc_b200._train import shard_rows

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
for dtype, tol in (("f32", 1e-5), ("bf16", 2e-2), ("f32x3", 1e-5)):
    J.set_compute_dtype(dtype)
    B, d = 512, 256
    g = torch.Generator().manual_seed(0)
    model = J.make_mlp(0, d, d, 4, d, "tanh", use_bias=True)          # same seed on every rank: replicated weights
    x = torch.randn(B, d, generator=g).cuda()
    y = torch.randn(B, d, generator=g).cuda()
    lo, hi = shard_rows(B, rank, world)
    opt = J.optim.sgd(0.1)
    kw = dict(ode_solver=J.Euler(), max_t1=5, dt=0.1 * B, stepsize_controller=J.ConstantStepSize())
    out = J.make_pc_step(model, opt, opt.init((model, None)), y[lo:hi].contiguous(), input=x[lo:hi].contiguous(), **kw)
    if rank == 0:
        dist.barrier()
    else:
        dist.barrier()
    # single-process reference on rank 0 (temporarily pretend there is no process group)
    if rank == 0:
        import jpc_b200._train as T
        saved = T._world
        T._world = lambda: 1
        real_all_reduce = T._Arena.all_reduce
        T._Arena.all_reduce = lambda self: None
        ref = J.make_pc_step(model, opt, opt.init((model, None)), y, input=x, **kw)
        T._world, T._Arena.all_reduce = saved, real_all_reduce
        err = max(float((a[1].weight - b[1].weight).abs().max() / b[1].weight.abs().max()) for a, b in zip(out["model"], ref["model"]))
        e_err = float((out["energies"] - ref["energies"]).abs().max() / ref["energies"].abs().max())
        l_err = abs(float(out["loss"]) - float(ref["loss"])) / abs(float(ref["loss"]))
        print(f"{dtype}: world={world} updated-weight rel err {err:.2e}, energies {e_err:.2e}, loss {l_err:.2e} (tol {tol:g})", flush=True)
        assert err < tol and e_err < max(tol, 1e-5) and l_err < max(tol, 1e-5)
dist.barrier()
if rank == 0:
    print("DP NCCL parity OK", flush=True)
dist.destroy_process_group()
