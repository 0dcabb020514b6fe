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
# ---- the reference's DEFAULT solve (Heun + PIDController, dt0 = None) under data parallelism: every rank must take the same
# accept / reject decisions (global norms, global 1/B in the initial-step heuristic) and reproduce the single-process result
import jpc_b200._train as T
J.set_compute_dtype("f32")
B, d = 97, 64                               # uneven shards: 49 + 48 rows
g = torch.Generator().manual_seed(1)
model = J.make_mlp(1, d, d, 3, 16, "tanh", use_bias=True)
x = torch.randn(B, d, generator=g).cuda()
y = torch.randn(B, 16, generator=g).cuda()
lo, hi = shard_rows(B, rank, world)
opt = J.optim.sgd(0.1)
T._GLOBAL_ROWS.clear()
out = J.make_pc_step(model, opt, opt.init((model, None)), y[lo:hi].contiguous(), input=x[lo:hi].contiguous(), max_t1=6)
stats = dict(J.solve_inference.last_stats)
gathered = [None] * world
dist.all_gather_object(gathered, (stats["accepted"], stats["rejected"]))
assert all(gq == gathered[0] for gq in gathered), f"ranks took different controller decisions: {gathered}"
if rank == 0:
    saved, real = T._world, T._Arena.all_reduce
    T._world = lambda: 1
    T._Arena.all_reduce = lambda self: None
    import jpc_b200._core as Cmod
    import torch.distributed as _d
    real_init = _d.is_initialized
    _d.is_initialized = lambda: False
    try:
        ref = J.make_pc_step(model, opt, opt.init((model, None)), y, input=x, max_t1=6)
        rstats = dict(J.solve_inference.last_stats)
    finally:
        _d.is_initialized = real_init
        T._world, T._Arena.all_reduce = saved, real
    err = max(float((a[1].weight - b[1].weight).abs().max() / b[1].weight.abs().max()) for a, b in zip(out["model"], ref["model"]))
    print(f"adaptive (Heun + PID, dt0=None), uneven shards {gathered}: single process {(rstats['accepted'], rstats['rejected'])}, "
          f"updated-weight rel err {err:.2e}", flush=True)
    assert (rstats["accepted"], rstats["rejected"]) == gathered[0] and err < 2e-5
dist.barrier()
# ---- bf16 gradient payload (bucketed exchange): within the bf16 path's 2e-2
J.set_compute_dtype("bf16")
J.set_grad_allreduce_dtype("bf16")
B, d = 512, 256
g = torch.Generator().manual_seed(0)
model = J.make_mlp(0, d, d, 4, d, "tanh", use_bias=True)
x = torch.randn(B, d, generator=g).cuda()
y = torch.randn(B, d, generator=g).cuda()
lo, hi = shard_rows(B, rank, world)
kw = dict(ode_solver=J.Euler(), max_t1=5, dt=0.1 * B, stepsize_controller=J.ConstantStepSize())
T._GLOBAL_ROWS.clear()
out = J.make_pc_step(model, opt, opt.init((model, None)), y[lo:hi].contiguous(), input=x[lo:hi].contiguous(), **kw)
J.set_grad_allreduce_dtype("f32")
out32 = J.make_pc_step(model, opt, opt.init((model, None)), y[lo:hi].contiguous(), input=x[lo:hi].contiguous(), **kw)
if rank == 0:
    err = max(float((a[1].weight - b[1].weight).abs().max() / b[1].weight.abs().max()) for a, b in zip(out["model"], out32["model"]))
    print(f"bf16 gradient payload vs fp32 payload: updated-weight rel err {err:.2e}", flush=True)
    assert err < 2e-2
dist.barrier()
# ---- PCTrainer under data parallelism: train step + captured NCCL all-reduce + SGD update as one replayed CUDA graph ----
for dtype, tol in (("f32", 1e-5), ("bf16", 2e-2)):
    J.set_compute_dtype(dtype)
    B, d = 256, 128
    g = torch.Generator().manual_seed(3)
    model = J.make_mlp(3, d, d, 4, d, "tanh", use_bias=True)
    lo, hi = shard_rows(B, rank, world)
    kw = dict(ode_solver=J.Euler(), max_t1=4, dt=0.5, stepsize_controller=J.ConstantStepSize())
    T._GLOBAL_ROWS.clear()
    tr = J.PCTrainer(model, J.optim.sgd(0.05), batch_size=hi - lo, **kw)
    opt = J.optim.sgd(0.05)
    ref_model, ref_st = model, opt.init((model, None))
    for it in range(3):
        x = torch.randn(B, d, generator=g).cuda()
        y = torch.randn(B, d, generator=g).cuda()
        out = tr.step(y[lo:hi].contiguous(), x[lo:hi].contiguous())
        dist.barrier()                      # an eager collective on the default communicator between graph replays
        if rank == 0:
            saved, real = T._world, T._Arena.all_reduce
            T._world = lambda: 1
            T._Arena.all_reduce = lambda self: None
            try:
                ref = J.make_pc_step(ref_model, opt, ref_st, y, input=x, **kw)
            finally:
                T._world, T._Arena.all_reduce = saved, real
            ref_model, ref_st = ref["model"], ref["opt_state"]
            err = max(float((a[1].weight - b[1].weight).abs().max() / b[1].weight.abs().max()) for a, b in zip(tr.model, ref_model))
            l_err = abs(float(out["loss"]) - float(ref["loss"])) / abs(float(ref["loss"]))
            print(f"PCTrainer[{dtype}] DP step {it}: weights vs single-process make_pc_step rel err {err:.2e}, loss {l_err:.2e}", flush=True)
            assert err < tol and l_err < max(tol, 1e-5)
dist.barrier()
if rank == 0:
    print("DP NCCL parity OK", flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)   # (see bench.py: no communicator teardown while a graph that captured a collective is alive)
