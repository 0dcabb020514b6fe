"""Launch plans (include/jpc_b200.h "launch plans", csrc/plan.cu): an entry point called again with identical arguments
replays a CUDA graph captured inside the library instead of re-enqueueing its kernels.  Results must not depend on it."""
import pytest
import torch

from tests.helpers import data, dev, f32_model, rel, to_gpu_model
from oracle import pc_oracle as O

pytestmark = pytest.mark.gpu


def _setup(dtype, dims, B, act="tanh", bias=True):
    import jpc_b200 as J
    J.set_compute_dtype(dtype)
    om = f32_model(O.make_mlp(0, dims[0], dims[1], dims[2], dims[3], act, use_bias=bias))
    x, y = data(B, dims[0], dims[3], onehot_out=False)
    return J, om, to_gpu_model(om), dev(x), dev(y), x.float().double(), y.float().double()


@pytest.mark.parametrize("dtype,tol,dims,B", [("f32", 1e-5, (24, 48, 4, 8), 16), ("bf16", 2e-2, (128, 256, 4, 64), 256)])
def test_plan_replay_of_the_fused_train_step_matches_the_oracle(dtype, tol, dims, B):
    """jpcb_pc_train_step on fixed buffers, six times: calls 1-2 enqueue kernels, call 3 captures, calls 4-6 replay.  Every
    call must give the oracle's step; the launch counter advances by the same amount per call either way."""
    from jpc_b200 import _lib
    from jpc_b200._core import _Net
    from jpc_b200._train import _Arena
    J, om, gm, x, y, xd, yd = _setup(dtype, dims, B)
    ref = O.pc_train_step(om, None, xd, yd, solver="euler", max_t1=3.0, dt=0.5)
    J.plan_cache_clear()
    J.plan_cache_configure(16, 3)
    s0 = J.plan_cache_stats()
    net = _Net(gm, None, x, y, param_type="sp", loss_id="mse", gamma=None, output_energy_scaling=None, activity_decay=0.0,
               weight_decay=0.0, spectral_penalty=0.0, solver=_lib.SOLVER_EULER, n_steps=6, dt=0.5, dt_last=0.5)
    arena, acts, ws = _Arena(net), net.new_acts(), net.workspace()
    per_call = []
    for it in range(6):
        arena.flat.zero_()
        for a in acts:
            a.zero_()
        l0 = _lib.lib.jpcb_launch_count()
        _lib.check(_lib.lib.jpcb_pc_train_step(net.desc, net.W, net.b, _lib.ptr(x), _lib.ptr(y), _lib.ptr_array(acts),
                                               _lib.ptr(arena.loss), _lib.ptr(arena.E), _lib.ptr_array(arena.gW),
                                               _lib.ptr_array(arena.gb), _lib.ptr(ws), ws.numel(), net.stream()))
        per_call.append(_lib.lib.jpcb_launch_count() - l0)
        torch.cuda.synchronize()
        assert abs(float(arena.loss) - float(ref["loss"])) <= tol * abs(float(ref["loss"])), it
        assert rel(arena.E, ref["energies"]) < tol, it
        for p, q in zip(acts[:-1], ref["activities"][:-1]):
            assert rel(p, q) < tol, it
        gW_ref = [g["W"] for g in ref["grads"]] if isinstance(ref["grads"][0], dict) else None
        if gW_ref is not None:
            gmax = max(float(g.abs().max()) for g in gW_ref)
            for p, q in zip(arena.gW, gW_ref):
                assert float((p.double().cpu() - q).abs().max()) <= tol * gmax, it
    s1 = J.plan_cache_stats()
    J.plan_cache_configure(16, 2)   # (back to the defaults)
    assert s1["captures"] - s0["captures"] == 1
    assert s1["replays"] - s0["replays"] == 3
    assert len(set(per_call)) == 1 and per_call[0] > 0, per_call   # replayed kernels are counted like enqueued ones


def test_plans_do_not_change_a_functional_training_loop():
    """Ten `make_pc_step` steps (fresh outputs every step: the pointer sets repeat as the caching allocator recycles
    blocks) with plans enabled and disabled, from the same start: same losses and same final parameters."""
    J, om, gm, x, y, _, _ = _setup("f32", (20, 32, 5, 6), 16)
    kw = dict(input=x, ode_solver=J.Euler(), max_t1=4, dt=0.5, stepsize_controller=J.ConstantStepSize())

    def loop(enabled):
        J.plan_cache_clear()
        J.plan_cache_configure(16 if enabled else 0, 2)
        opt = J.optim.sgd(0.05)
        model, st, losses = gm, opt.init((gm, None)), []
        for _ in range(10):
            out = J.make_pc_step(model, opt, st, y, **kw)
            model, st = out["model"], out["opt_state"]
            losses.append(float(out["loss"]))
        return losses, [p.clone() for p in J.nn.tree_leaves(model)], J.plan_cache_stats()
    try:
        la, pa, sa = loop(True)
        lb, pb, _ = loop(False)
    finally:
        J.plan_cache_configure(16, 2)
    print("plan cache over a 10-step functional loop:", sa)
    for a, b in zip(la, lb):
        assert abs(a - b) <= 1e-6 * abs(b)
    for a, b in zip(pa, pb):
        assert rel(a, b.double().cpu()) < 1e-6
    assert la[-1] < la[0]


def test_calls_under_the_callers_own_capture_bypass_the_plans():
    """PCTrainer captures jpcb_pc_train_step into ITS graph: the library must enqueue the kernels themselves there."""
    J, om, gm, x, y, _, _ = _setup("f32", (20, 32, 3, 6), 16)
    s0 = J.plan_cache_stats()
    tr = J.PCTrainer(gm, J.optim.sgd(0.05), batch_size=16, ode_solver=J.Euler(), max_t1=2, dt=0.5,
                     stepsize_controller=J.ConstantStepSize())
    out = tr.step(y, x)
    torch.cuda.synchronize()
    assert J.plan_cache_stats()["bypassed"] > s0["bypassed"]
    assert float(out["loss"]) > 0


def test_returned_model_carries_its_lowering_and_loses_it_when_touched():
    """`make_pc_step` returns its model as a list that remembers how it was lowered (layer forms, dims, parameter
    addresses): the next call trusts that only while every layer and parameter object is untouched."""
    J, om, gm, x, y, _, _ = _setup("f32", (20, 32, 5, 6), 16)
    kw = dict(input=x, ode_solver=J.Euler(), max_t1=2, dt=0.5, stepsize_controller=J.ConstantStepSize())
    opt = J.optim.adam(1e-2)
    out = J.make_pc_step(gm, opt, opt.init((gm, None)), y, **kw)
    m = out["model"]
    assert isinstance(m, list) and getattr(m, "_jpcb", None) is not None and len(m) == 5
    acts = J.init_activities_with_ffwd(m, x)
    e_memo = J.pc_energy_fn((m, None), acts, y, x=x)
    e_full = J.pc_energy_fn((list(m), None), acts, y, x=x)          # a plain copy: derived from scratch
    assert abs(float(e_memo) - float(e_full)) <= 1e-6 * abs(float(e_full))   # (energy sums use atomics: equal up to summation order)
    m[2].layers[1].weight = m[2].layers[1].weight * 2.0              # touch one parameter: the remembered lowering must not be used
    e_new = J.pc_energy_fn((m, None), acts, y, x=x)
    e_ref = J.pc_energy_fn((list(m), None), acts, y, x=x)
    assert abs(float(e_new) - float(e_ref)) <= 1e-6 * abs(float(e_ref)) and abs(float(e_new) - float(e_memo)) > 1e-3 * abs(float(e_memo))
    out2 = J.make_pc_step(out["model"], opt, out["opt_state"], y, **kw)   # and training simply goes on from the touched model
    assert float(out2["loss"]) > 0 and out2["model"]._jpcb is not None


def test_plan_eviction_and_capture_budget_keep_results_right():
    """Two plans of room, three call signatures used round-robin: plans are evicted (and their graphs destroyed only after
    their last launch completed) and re-captured; past the capture budget the calls simply enqueue their kernels.  Every call
    must still give the same numbers as with plans off."""
    from jpc_b200 import _lib
    from jpc_b200._core import _Net
    from jpc_b200._train import _Arena
    J, om, gm, x, y, _, _ = _setup("f32", (24, 48, 4, 8), 16)
    net = _Net(gm, None, x, y, param_type="sp", loss_id="mse", gamma=None, output_energy_scaling=None, activity_decay=0.0,
               weight_decay=0.0, spectral_penalty=0.0, solver=_lib.SOLVER_EULER, n_steps=3, dt=0.5, dt_last=0.5)
    ws = net.workspace()
    sets = [(_Arena(net), net.new_acts()) for _ in range(3)]

    def call(k):
        arena, acts = sets[k]
        arena.flat.zero_()
        _lib.check(_lib.lib.jpcb_pc_train_step(net.desc, net.W, net.b, _lib.ptr(x), _lib.ptr(y), _lib.ptr_array(acts),
                                               _lib.ptr(arena.loss), _lib.ptr(arena.E), _lib.ptr_array(arena.gW),
                                               _lib.ptr_array(arena.gb), _lib.ptr(ws), ws.numel(), net.stream()))
        torch.cuda.synchronize()
        return arena.flat.clone()
    try:
        J.plan_cache_clear()
        J.plan_cache_configure(0, 2)
        want = call(0)
        J.plan_cache_configure(2, 2)
        s0 = J.plan_cache_stats()
        for it in range(30):
            got = call(it % 3)
            assert rel(got, want.double().cpu()) < 1e-6, it
        s1 = J.plan_cache_stats()
    finally:
        J.plan_cache_clear()
        J.plan_cache_configure(16, 2)
    print("eviction run:", {k: s1[k] - s0[k] for k in s1})
    assert s1["captures"] - s0["captures"] >= 3          # more signatures than room: something was captured again
    assert s1["captures"] - s0["captures"] <= 12         # ... but the budget stopped the thrashing


@pytest.mark.parametrize("dtype,tol,dims,B", [("f32", 1e-5, (64, 256, 4, 32), 512), ("bf16", 2e-2, (128, 256, 4, 64), 512),
                                              ("f32", 1e-5, (33, 65, 7, 9), 37)])
def test_reproducible_sums_are_bitwise_stable_and_right(dtype, tol, dims, B):
    """jpcb_set_reproducible: loss and energies of the fused train step by fixed-order sums over the stored errors -- the same
    bits every run, the oracle's values, and (within summation-order noise) what the default atomics give."""
    from jpc_b200 import _lib
    from jpc_b200._core import _Net
    from jpc_b200._train import _Arena
    J, om, gm, x, y, xd, yd = _setup(dtype, dims, B)
    ref = O.pc_train_step(om, None, xd, yd, solver="euler", max_t1=2.0, dt=0.5)
    net = _Net(gm, None, x, y, param_type="sp", loss_id="mse", gamma=None, output_energy_scaling=None, activity_decay=0.0,
               weight_decay=0.0, spectral_penalty=0.0, solver=_lib.SOLVER_EULER, n_steps=4, dt=0.5, dt_last=0.5)

    def run():
        arena, acts, ws = _Arena(net), net.new_acts(), net.workspace()
        arena.flat.zero_()
        gW, gb = arena.grad_ptr_arrays()
        _lib.check(_lib.lib.jpcb_pc_train_step(net.desc, net.W, net.b, _lib.ptr(x), _lib.ptr(y), _lib.ptr_array(acts),
                                               _lib.ptr(arena.loss), _lib.ptr(arena.E), gW, gb, _lib.ptr(ws), ws.numel(), net.stream()))
        torch.cuda.synchronize()
        return arena.E.clone(), arena.loss.clone()
    try:
        J.plan_cache_configure(0, 2)       # (every call enqueues its kernels: block scheduling is free to differ)
        J.set_reproducible(True)
        runs = [run() for _ in range(4)]
        J.set_reproducible(False)
        E_at, loss_at = run()
    finally:
        J.set_reproducible(False)
        J.plan_cache_configure(16, 2)
    for E, loss in runs[1:]:
        assert torch.equal(E, runs[0][0]) and torch.equal(loss, runs[0][1])
    E, loss = runs[0]
    assert abs(float(loss) - float(ref["loss"])) <= tol * abs(float(ref["loss"]))
    assert rel(E, ref["energies"]) < tol
    assert rel(E, E_at.double().cpu()) < (1e-5 if dtype == "f32" else 5e-3)   # (bf16 path: the fixed-order sums read bf16 error copies)
    assert abs(float(loss) - float(loss_at)) <= 1e-5 * abs(float(loss_at))
