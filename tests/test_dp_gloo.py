"""Data-parallel host logic on CPU: world_size-2 `gloo` processes, the oracle standing in for the
CUDA kernels.  Checks what the N>1 path of jpc_b200.make_pc_step relies on: contiguous
row shards, every rank normalising by the GLOBAL batch, ONE all-reduce (SUM) of the flat arena
[gW, gb, E_layers, loss] reproducing the full-batch train step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pc_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from jpc_b200._train import _Arena, shard_rows
        B, dims = 11, (6, 12, 3, 4)  # 11 rows over 2 ranks: a ragged 6 + 5 split
        om = O.make_mlp(0, dims[0], dims[1], dims[2], dims[3], "tanh", use_bias=True)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, dims[0], generator=g, dtype=torch.float64)
        y = torch.randn(B, dims[3], generator=g, dtype=torch.float64)
        lo, hi = shard_rows(B, rank, world)
        scale = (hi - lo) / B  # local results are normalised by the local batch in the oracle; the kernels use B_global
        # dt enters the dynamics as dt / B: keep the per-sample step and the step count of the full-batch run
        r = O.pc_train_step(om, None, x[lo:hi], y[lo:hi], solver="heun", max_t1=3.0 * (hi - lo) / B, dt=0.5 * (hi - lo) / B)
        arena = _Arena(weight_shapes=[tuple(l["W"].shape) for l in om], bias_shapes=[tuple(l["b"].shape) for l in om],
                       device="cpu")
        for l, gr in enumerate(r["grads"]):
            arena.gW[l].copy_((gr["W"] * scale).float())
            arena.gb[l].copy_((gr["b"] * scale).float())
        arena.E.copy_((r["energies"] * scale).float())
        arena.loss.copy_((r["loss"] * scale).float())
        arena.all_reduce()
        if rank == 0:
            full = O.pc_train_step(om, None, x, y, solver="heun", max_t1=3.0, dt=0.5)
            err = max(O.max_rel_err(arena.gW[l], full["grads"][l]["W"]) for l in range(len(om)))
            err = max(err, max(O.max_rel_err(arena.gb[l], full["grads"][l]["b"]) for l in range(len(om))))
            err = max(err, O.max_rel_err(arena.E, full["energies"]), abs(float(arena.loss) - float(full["loss"])) / float(full["loss"]))
            q.put(err)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() < 1e-5  # fp32 arena


def _bpc_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace
        from jpc_b200._bpc import _grad_arena
        from jpc_b200._train import shard_rows
        B, H, dims = 9, 2, (5, 8, 8, 3)  # 9 rows over 2 ranks: 5 + 4
        td = O.make_mlp(0, dims[0], dims[1], H + 1, dims[3], "tanh", use_bias=True)
        bu = O.make_mlp(1, dims[3], dims[1], H + 1, dims[0], "tanh", use_bias=True)[::-1]
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, dims[0], generator=g, dtype=torch.float64)
        y = torch.randn(B, dims[3], generator=g, dtype=torch.float64)
        acts = [a + 0.3 * torch.randn(a.shape, generator=g, dtype=torch.float64) for a in O.init_activities_with_ffwd(td, x)]
        lo, hi = shard_rows(B, rank, world)
        scale = (hi - lo) / B   # the oracle normalises by the rows it sees; the kernels by desc.batch_global
        loc = [a[lo:hi] for a in acts]
        tdg, bug = O.compute_bpc_param_grads(td, bu, loc, y[lo:hi], x=x[lo:hi])
        E = O.bpc_energy_fn(td, bu, loc, y[lo:hi], x=x[lo:hi], record_layers=True)
        net = SimpleNamespace(dims=list(dims), H=H, td_bias=True, bu_bias=True, dev=torch.device("cpu"))
        flat, gW, gbW, gV, gbV, Ev = _grad_arena(net, True, zero=True)
        for j in range(H + 1):
            gW[j].copy_((tdg[j]["W"] * scale).float()); gbW[j].copy_((tdg[j]["b"] * scale).float())
            gV[j].copy_((bug[j]["W"] * scale).float()); gbV[j].copy_((bug[j]["b"] * scale).float())
        Ev.copy_((E * scale).float())
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)   # what _param_grads(exchange=True) does
        if rank == 0:
            ftd, fbu = O.compute_bpc_param_grads(td, bu, acts, y, x=x)
            fE = O.bpc_energy_fn(td, bu, acts, y, x=x, record_layers=True)
            err = O.max_rel_err(Ev, fE)
            for j in range(H + 1):
                err = max(err, O.max_rel_err(gW[j], ftd[j]["W"]), O.max_rel_err(gbW[j], ftd[j]["b"]),
                          O.max_rel_err(gV[j], fbu[j]["W"]), O.max_rel_err(gbV[j], fbu[j]["b"]))
            q.put(err)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_bpc_gradient_exchange_matches_full_batch():
    """bPC under data parallelism: the flat buffer `bpc_train_step` all-reduces (both models' gradients + energy pairs, each
    rank's part normalised by the global batch) sums to the full-batch values; ragged 5 + 4 row split."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_bpc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() < 1e-5  # fp32 buffer


def test_shard_rows_cover_and_are_contiguous():
    from jpc_b200._train import shard_rows
    for n, w in ((10, 2), (10, 3), (4096, 8), (7, 8), (1, 1)):
        spans = [shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_arena_layout():
    from jpc_b200._train import _Arena
    a = _Arena(weight_shapes=[(5, 3), (2, 5)], bias_shapes=[(5,), (2,)], device="cpu")
    ptrs = [t.data_ptr() for t in a.gW + a.gb + [a.E, a.loss]]
    assert all(p % 256 == 0 for p in [q - a.flat.data_ptr() for q in ptrs])
    assert a.E.shape == (2,) and a.loss.shape == ()
    a.flat.zero_()
    a.gW[1].fill_(3.0)
    assert float(a.flat.sum()) == 30.0


def test_arena_buckets_cover_the_arena():
    """The bucketed exchange (jpc_b200._train.dp_param_grads_and_exchange) reduces `flat[lo:hi]` per layer group: the
    groups must be contiguous, ordered, cover every layer once and the whole flat buffer (tail [E, loss] in the last)."""
    from jpc_b200._train import _Arena
    for L, bias in ((8, False), (3, True), (1, False), (13, True)):
        ws = [(16 + 8 * l, 24 + 4 * l) for l in range(L)]
        bs = [(16 + 8 * l,) for l in range(L)] if bias else None
        arena = _Arena(weight_shapes=ws, bias_shapes=bs, device="cpu")
        for nb in (1, 2, 4, 7, 50):
            bk = arena.buckets(nb)
            assert bk[0][0] == 0 and bk[0][2] == 0 and bk[-1][1] == L and bk[-1][3] == arena.flat.numel()
            assert len(bk) <= min(nb, L)
            for (a0, a1, f0, f1), (b0, b1, g0, g1) in zip(bk[:-1], bk[1:]):
                assert a1 == b0 and f1 == g0 and a0 < a1 and f0 < f1
            # every gradient view lies inside the flat range of its layer's bucket
            for llo, lhi, flo, fhi in bk:
                for l in range(llo, lhi):
                    off = arena.gW[l].storage_offset()
                    assert flo <= off and off + arena.gW[l].numel() <= fhi
            assert arena.E.storage_offset() >= bk[-1][2]
