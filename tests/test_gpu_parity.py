"""GPU parity tests: the CUDA path (through the C-ABI, via the reference-shaped Python API) against
the fp64 oracle on the same seeded inputs.  Tolerances are north_star's: 1e-5 max-norm relative
error in fp32, 2e-2 in bf16 (SURVEY 8d 'parity gates')."""
import math
import os

import pytest
import torch

from oracle import pc_oracle as O
from tests.helpers import GROUP_ARM, N_GATED, NOISE_ARM, data, dev, devs, f32_model, gate_all, perturb, rel, to_gpu_model

pytestmark = pytest.mark.gpu

TOL32 = 1e-5
TOLBF = 2e-2
GROUP_ARM_BUDGET = 400  # measured: 324 of 1215 gated tensors (profiles/r02_parity_gates.txt), nearly all in config 3 (see test_zz_group_arm_budget)


@pytest.fixture(autouse=True)
def _f32():
    import jpc_b200 as J
    J.set_compute_dtype("f32")
    yield
    J.set_compute_dtype("f32")


def _case(d_in, width, depth, d_out, act, B, use_bias=False, param_type="sp", skips=False, seed=0, **dkw):
    om = f32_model(O.make_mlp(seed, d_in, width, depth, d_out, act, use_bias=use_bias, param_type=param_type))
    sk = O.make_skip_model(depth) if skips else None
    x, y = data(B, d_in, d_out, **dkw)
    x, y = x.float().double(), y.float().double()
    return om, sk, x, y


def _gpu_skip(sk):
    import jpc_b200 as J
    return None if sk is None else J.make_skip_model(len(sk))


def _check_all(om, sk, x, y, tol, *, param_type="sp", gamma=None, lam=None, ad=0.0, perturbed=True):
    import jpc_b200 as J
    gm, gsk = to_gpu_model(om), _gpu_skip(sk)
    gx, gy = dev(x), dev(y)
    kw = dict(param_type=param_type, gamma=gamma, output_energy_scaling=lam)
    # feed-forward init
    a0 = O.init_activities_with_ffwd(om, x, skips=sk, param_type=param_type, gamma=gamma)
    ga0 = J.init_activities_with_ffwd(gm, gx, skip_model=gsk, param_type=param_type, gamma=gamma)
    for p, q in zip(ga0, a0):
        assert rel(p, q) < tol
    acts = [a.float().double() for a in (perturb(a0) if perturbed else a0)]
    gacts = devs(acts)
    # energies
    E = O.pc_energy_fn(om, sk, acts, y, x=x, record_layers=True, activity_decay=ad, **kw)
    gE = J.pc_energy_fn((gm, gsk), gacts, gy, x=gx, record_layers=True, activity_decay=ad, **kw)
    assert gE.shape == (len(om),)
    assert rel(gE, E) < tol
    F, g = O.compute_pc_activity_grad(om, sk, acts, y, x=x, activity_decay=ad, **kw)
    gF, gg = J.compute_pc_activity_grad((gm, gsk), gacts, gy, x=gx, activity_decay=ad, **kw)
    assert abs(float(gF) - float(F)) <= tol * abs(float(F))
    assert abs(float(J.pc_energy_fn((gm, gsk), gacts, gy, x=gx, activity_decay=ad, **kw)) - float(F)) <= tol * abs(float(F))
    gate_all(gg, g, tol, "check_all.activity_grads")
    # parameter gradients
    pg = O.compute_pc_param_grads(om, sk, acts, y, x=x, activity_decay=ad, **kw)
    gpg, gskip = J.compute_pc_param_grads((gm, gsk), gacts, gy, x=gx, activity_decay=ad, **kw)
    assert (gskip is None) == (gsk is None)
    for l, q in enumerate(pg):
        assert rel(gpg[l][1].weight, q["W"]) < tol
        if q["b"] is not None:
            assert rel(gpg[l][1].bias, q["b"]) < tol


@pytest.mark.parametrize("act", O.ACT_FNS)
@pytest.mark.parametrize("use_bias", [False, True])
def test_fixture_shapes_all_activations(act, use_bias):
    """reference tests/conftest.py shapes: batch 4, 10 -> 20 -> 20 -> 5."""
    _check_all(*_case(10, 20, 3, 5, act, 4, use_bias=use_bias), TOL32)


@pytest.mark.parametrize("param_type", ["mupc", "ntp"])
def test_skips_scalings_gamma_decay(param_type):
    om, sk, x, y = _case(12, 16, 6, 3, "relu", 5, param_type=param_type, skips=True)
    _check_all(om, sk, x, y, TOL32, param_type=param_type, gamma=0.5, lam=2.5, ad=0.1)


def test_ragged_dims_not_multiple_of_4():
    _check_all(*_case(7, 13, 4, 3, "tanh", 3, use_bias=True), TOL32)
    _check_all(*_case(33, 65, 3, 1, "gelu", 129, use_bias=True), TOL32)


def test_two_layer_minimum():
    _check_all(*_case(6, 9, 2, 4, "silu", 2), TOL32)


def _solve_case(om, sk, x, y, solver, max_t1, dt, tol, param_type="sp", ad=0.0, perturbed=False):
    import jpc_b200 as J
    gm, gsk = to_gpu_model(om), _gpu_skip(sk)
    gx, gy = dev(x), dev(y)
    a0 = O.init_activities_with_ffwd(om, x, skips=sk, param_type=param_type)
    if perturbed:
        a0 = [a.float().double() for a in perturb(a0)]
    sol, n = O.solve_inference(om, sk, a0, y, x=x, solver=solver, max_t1=max_t1, dt=dt, param_type=param_type,
                               activity_decay=ad, use_closed_form=True, event_atol=0.0, event_rtol=0.0)
    gsol = J.solve_inference((gm, gsk), devs(a0), gy, input=gx, param_type=param_type,
                             solver=J.Euler() if solver == "euler" else J.Heun(), max_t1=max_t1, dt=dt,
                             stepsize_controller=J.ConstantStepSize(), activity_decay=ad)
    assert len(gsol) == len(om)
    for p, q in zip(gsol, sol):
        assert p.ndim == 3 and p.shape[0] == 1
        assert rel(p[0], q) < tol
    return n


@pytest.mark.parametrize("solver", ["euler", "heun"])
def test_solve_inference_small(solver):
    om, sk, x, y = _case(10, 20, 3, 5, "tanh", 8, use_bias=True)
    assert _solve_case(om, sk, x, y, solver, 5.0, 0.5, TOL32) == 10
    # non-binary dt: the last step is clipped to t1
    assert _solve_case(om, sk, x, y, solver, 1.0, 0.3, TOL32, perturbed=True) == 4
    # stress step size ~ 0.1 * B
    _solve_case(om, sk, x, y, solver, 8.0, 0.8, TOL32, perturbed=True, ad=0.05)


def test_euler_equals_looped_update_pc_activities():
    """Euler(dt) == optax.sgd(dt) applied T times (SURVEY 8c iii)."""
    import jpc_b200 as J
    om, sk, x, y = _case(10, 20, 4, 5, "relu", 8)
    gm = to_gpu_model(om)
    gx, gy = dev(x), dev(y)
    acts = J.init_activities_with_ffwd(gm, gx)
    sol = J.solve_inference((gm, None), acts, gy, input=gx, solver=J.Euler(), max_t1=3.0, dt=0.5,
                            stepsize_controller=J.ConstantStepSize())
    opt = J.optim.sgd(0.5)
    st = opt.init(acts)
    a = acts
    for _ in range(6):
        r = J.update_pc_activities((gm, None), a, opt, st, gy, input=gx)
        assert set(r) == {"energy", "activities", "grads", "opt_state"}
        a, st = r["activities"], r["opt_state"]
    for p, q in zip(sol, a):
        assert rel(p[0], q) < 1e-6


def _train_case(om, sk, x, y, solver, max_t1, dt, tol, param_type="sp", zero_tol=1e-12):
    import jpc_b200 as J
    gm, gsk = to_gpu_model(om), _gpu_skip(sk)
    gx, gy = dev(x), dev(y)
    ref = O.pc_train_step(om, sk, x, y, solver=solver, max_t1=max_t1, dt=dt, param_type=param_type)
    opt = J.optim.sgd(0.0)
    out = J.make_pc_step(gm, opt, opt.init((gm, gsk)), gy, input=gx, param_type=param_type,
                         ode_solver=J.Euler() if solver == "euler" else J.Heun(), max_t1=max_t1, dt=dt,
                         stepsize_controller=J.ConstantStepSize(), skip_model=gsk, grad_norms=True)
    assert set(out) == {"model", "skip_model", "opt_state", "loss", "acc", "activities", "t_max", "energies",
                        "activity_norms", "model_param_norms", "skip_model_param_norms", "model_grad_norms",
                        "skip_model_grad_norms"}
    assert abs(float(out["loss"]) - float(ref["loss"])) <= tol * abs(float(ref["loss"]))
    assert rel(out["energies"], ref["energies"]) < tol
    gate_all([p[0] for p in out["activities"]], ref["activities"], tol, "train_case.activities")
    return out, ref


def _check_train_grads(om, sk, x, y, solver, max_t1, dt, tol, param_type="sp"):
    """Parameter gradients of the fused train step (read back through update_pc_params-free path)."""
    import jpc_b200 as J
    from jpc_b200._train import _Arena  # noqa: F401
    gm, gsk = to_gpu_model(om), _gpu_skip(sk)
    gx, gy = dev(x), dev(y)
    ref = O.pc_train_step(om, sk, x, y, solver=solver, max_t1=max_t1, dt=dt, param_type=param_type)
    # sgd(lr=1): new = old - g  =>  g = old - new
    opt = J.optim.sgd(1.0)
    out = J.make_pc_step(gm, opt, opt.init((gm, gsk)), gy, input=gx, param_type=param_type,
                         ode_solver=J.Euler() if solver == "euler" else J.Heun(), max_t1=max_t1, dt=dt,
                         stepsize_controller=J.ConstantStepSize(), skip_model=gsk)
    gmax = max(float(g["W"].abs().max()) for g in ref["grads"])
    for l, q in enumerate(ref["grads"]):
        g = gm[l][1].weight.double() - out["model"][l][1].weight.double()
        # the subtraction itself costs ~1e-7 * |W|; compare on the gradient's own scale
        wmax = float(gm[l][1].weight.abs().max())
        assert float((g.cpu() - q["W"]).abs().max()) < tol * max(float(q["W"].abs().max()), 1e-30) + 2e-7 * wmax or \
            float((g.cpu() - q["W"]).abs().max()) < tol * gmax


def test_make_pc_step_small():
    om, sk, x, y = _case(10, 20, 3, 5, "relu", 4)
    _train_case(om, sk, x, y, "euler", 5.0, 0.5, TOL32)
    _train_case(om, sk, x, y, "heun", 5.0, 0.5, TOL32)


def _fused_step_raw(om, sk, x, y, solver, max_t1, dt, param_type="sp"):
    """jpcb_pc_train_step through the Python plumbing, returning raw gradients."""
    import jpc_b200 as J
    from jpc_b200 import _lib
    from jpc_b200._core import _Net, _solver_id, _time_grid
    from jpc_b200._train import _Arena
    gm, gsk = to_gpu_model(om), _gpu_skip(sk)
    gx, gy = dev(x), dev(y)
    T, h, hl = _time_grid(max_t1, dt)
    net = _Net(gm, gsk, gx, gy, param_type=param_type, loss_id="mse", gamma=None, output_energy_scaling=None,
               activity_decay=0.0, weight_decay=0.0, spectral_penalty=0.0,
               solver=_solver_id(J.Euler() if solver == "euler" else J.Heun()), n_steps=T, dt=h, dt_last=hl)
    arena = _Arena(net)
    acts = net.new_acts()
    ws = net.workspace()
    _lib.check(_lib.lib.jpcb_pc_train_step(net.desc, net.W, net.b, _lib.ptr(net.x), _lib.ptr(net.y), _lib.ptr_array(acts),
                                           _lib.ptr(arena.loss), _lib.ptr(arena.E), _lib.ptr_array(arena.gW),
                                           _lib.ptr_array(arena.gb) if arena.gb else None, _lib.ptr(ws), ws.numel(),
                                           net.stream()))
    torch.cuda.synchronize()
    arena.gW, arena.gb = net.out_grads(arena.gW, arena.gb)  # (slices the zero padding of the bf16 path off again)
    return net.out_acts(acts), arena


def _check_fused(om, sk, x, y, solver, max_t1, dt, tol, param_type="sp", kink_budget=0):
    """`kink_budget`: how many activity entries may miss `tol` (by at most 200x).  relu' is discontinuous: an fp32
    pre-activation within rounding distance of 0 can take the other branch for one evaluation, which moves that one
    entry by ~1e-4 relative.  Over config 2's 1.3e5 activity entries x 100 evaluations about one such event is
    expected; which entry it hits depends on the summation order of the kernel (measured: scripts/warp8_check.py --
    one entry out of 131 072 with one kernel, none with the other, none with tanh)."""
    ref = O.pc_train_step(om, sk, x, y, solver=solver, max_t1=max_t1, dt=dt, param_type=param_type)
    acts, arena = _fused_step_raw(om, sk, x, y, solver, max_t1, dt, param_type)
    assert abs(float(arena.loss) - float(ref["loss"])) <= tol * abs(float(ref["loss"]))
    # every layer's energy on its own scale (a layer whose energy is exactly zero in the oracle must be ~0)
    gate_all([e for e in arena.E], [e for e in ref["energies"]], tol, "check_fused.energies", zero_abs=1e-10)
    if kink_budget:
        amax = max(float(t.abs().max()) for t in ref["activities"])
        for p, q in zip(acts, ref["activities"]):
            d = (p.double().cpu() - q).abs()
            lim = tol * max(float(q.abs().max()), amax)
            assert int((d > lim).sum()) <= kink_budget and float(d.max()) < 200 * lim
        # (a relu-kink entry moves its neighbours' gradients too: the gates below carry the same factor)
    gtol = tol
    if not kink_budget:
        gate_all(acts, ref["activities"], tol, "check_fused.activities")
    gate_all(arena.gW, [q["W"] for q in ref["grads"]], gtol, "check_fused.gW")
    if ref["grads"][0]["b"] is not None:
        gate_all(arena.gb, [q["b"] for q in ref["grads"]], gtol, "check_fused.gb")


def test_config1_discriminative_full_size():
    """BASELINE config 1: 784-300-300-10 tanh, batch 64, 20 Euler steps, fp32."""
    om, sk, x, y = _case(784, 300, 3, 10, "tanh", 64)
    _check_fused(om, sk, x, y, "euler", 10.0, 0.5, TOL32)
    _check_all(om, sk, x, y, TOL32)


def test_config2_generative_full_size():
    """BASELINE config 2: 10-256-256-784 relu+bias, batch 256, 50 Heun steps."""
    om, sk, x, y = _case(10, 256, 3, 784, "relu", 256, use_bias=True, onehot_in=True, onehot_out=False)
    _check_fused(om, sk, x, y, "heun", 25.0, 0.5, TOL32, kink_budget=3)


def test_config3_mupc_deep_residual_full_size():
    """BASELINE config 3: muPC width 128 depth 128 + identity skips, batch 128, 20 steps."""
    om, sk, x, y = _case(784, 128, 128, 10, "relu", 128, param_type="mupc", skips=True)
    _check_fused(om, sk, x, y, "euler", 10.0, 0.5, TOL32, param_type="mupc")
    _check_all(om, sk, x, y, TOL32, param_type="mupc")


# ---- bf16 tensor-core path -------------------------------------------------------------------
def test_bf16_path_medium():
    import jpc_b200 as J
    J.set_compute_dtype("bf16")
    om, sk, x, y = _case(256, 512, 4, 128, "tanh", 384, use_bias=True, onehot_out=False)
    _check_all(om, sk, x, y, TOLBF)
    _check_fused(om, sk, x, y, "euler", 5.0, 0.5, TOLBF)
    _check_fused(om, sk, x, y, "heun", 2.0, 0.5, TOLBF)


def test_bf16_path_ragged_tiles():
    """dims that are multiples of 8 but not of the 128x256x64 tile."""
    import jpc_b200 as J
    J.set_compute_dtype("bf16")
    om, sk, x, y = _case(72, 200, 3, 40, "relu", 136, onehot_out=False)
    _check_all(om, sk, x, y, TOLBF)
    _check_fused(om, sk, x, y, "euler", 2.0, 0.5, TOLBF)


def test_bf16_rejects_unaligned():
    import jpc_b200 as J
    J.set_compute_dtype("bf16")
    om, sk, x, y = _case(10, 20, 3, 5, "relu", 4)
    with pytest.raises(NotImplementedError):
        J.init_activities_with_ffwd(to_gpu_model(om), dev(x))


# ---- error behaviour (reference raises the same exception types) -----------------------------
def test_error_behaviour():
    import jpc_b200 as J
    om, sk, x, y = _case(10, 20, 3, 5, "relu", 4)
    gm = to_gpu_model(om)
    gx, gy = dev(x), dev(y)
    acts = J.init_activities_with_ffwd(gm, gx)
    with pytest.raises(ValueError):
        J.pc_energy_fn((gm, None), acts, gy, x=gx, param_type="bogus")
    with pytest.raises(ValueError):
        J.pc_energy_fn((gm, None), acts, gy, x=gx, loss="bogus")
    opt = J.optim.sgd(0.1)
    with pytest.raises(ValueError):
        J.make_pc_step(gm, opt, opt.init((gm, None)), gy, input=None)
    with pytest.raises(RuntimeError):  # no CPU path
        J.init_activities_with_ffwd(to_gpu_model(om, device="cpu"), x.float())


# ---- bidirectional PC (BASELINE config 5) ------------------------------------------------------
def _bpc_case(d_x, width, depth, d_y, act, B, use_bias=False, skips=False, seed=0):
    """top-down make_mlp(d_x -> d_y), bottom-up make_mlp(d_y -> d_x)[::-1] (examples/bidirectional_pc.ipynb cell 8)."""
    td = f32_model(O.make_mlp(seed, d_x, width, depth, d_y, act, use_bias=use_bias))
    bu = f32_model(O.make_mlp(seed + 1, d_y, width, depth, d_x, act, use_bias=use_bias))[::-1]
    sk = O.make_skip_model(depth) if skips else None
    x, y = data(B, d_x, d_y, onehot_in=True, onehot_out=False)
    x, y = x.float().double(), y.float().double()
    # activities z_0..z_{H-1} + unused slot: a forward sweep of the top-down model, perturbed so every error is active
    a0 = O.init_activities_with_ffwd(td, x)
    acts = [a.float().double() for a in perturb(a0)]
    return td, bu, sk, x, y, acts


def _check_bpc(td, bu, sk, x, y, acts, tol, af=1.0, ab=1.0):
    import jpc_b200 as J
    gtd, gbu, gsk = to_gpu_model(td), to_gpu_model(bu), _gpu_skip(sk)
    gx, gy, gacts = (None if x is None else dev(x)), dev(y), devs(acts)
    kw = dict(forward_energy_weight=af, backward_energy_weight=ab)
    E = O.bpc_energy_fn(td, bu, acts, y, x=x, skips=sk, record_layers=True, **kw)
    gE = J.bpc_energy_fn(gtd, gbu, gacts, gy, x=gx, skip_model=gsk, record_layers=True, **kw)
    assert gE.shape == (2 * len(td),)
    assert rel(gE, E) < tol
    F, g = O.compute_bpc_activity_grad(td, bu, acts, y, x=x, skips=sk, **kw)
    gF, gg = J.compute_bpc_activity_grad(gtd, gbu, gacts, gy, x=gx, skip_model=gsk, **kw)
    assert abs(float(gF) - float(F)) <= tol * abs(float(F))
    assert abs(float(J.bpc_energy_fn(gtd, gbu, gacts, gy, x=gx, skip_model=gsk, **kw)) - float(F)) <= tol * abs(float(F))
    gate_all(gg, g, tol, "check_bpc.activity_grads")
    tdg, bug = O.compute_bpc_param_grads(td, bu, acts, y, x=x, skips=sk, **kw)
    gtdg, gbug = J.compute_bpc_param_grads(gtd, gbu, gacts, gy, x=gx, skip_model=gsk, **kw)
    for got, want in ((gtdg, tdg), (gbug, bug)):
        for l, q in enumerate(want):
            assert rel(got[l][1].weight, q["W"]) < tol
            if q["b"] is not None:
                assert rel(got[l][1].bias, q["b"]) < tol


def _check_bpc_infer(td, bu, sk, x, y, acts, tol, T, dt, af=1.0, ab=1.0, kink_budget=0):
    """fused T Euler steps == T x (oracle gradient step with sgd(dt)).  `kink_budget`: activity entries per tensor allowed
    outside `tol` -- with relu, 1M entries x 30 steps, an fp32 pre-activation within rounding of 0 occasionally takes the
    other branch for one step (scripts/x3_check.py: 1 entry of 1 048 576, all others at 1e-8)."""
    import jpc_b200 as J
    kw = dict(forward_energy_weight=af, backward_energy_weight=ab)
    a = [t.clone() for t in acts]
    for _ in range(T):
        _, g = O.compute_bpc_activity_grad(td, bu, a, y, x=x, skips=sk, **kw)
        a = [p - dt * q for p, q in zip(a, g)]
    gtd, gbu, gsk = to_gpu_model(td), to_gpu_model(bu), _gpu_skip(sk)
    gx = None if x is None else dev(x)
    ga = J.solve_bpc_inference(gtd, gbu, devs(acts), dev(y), input=gx, n_steps=T, dt=dt, skip_model=gsk, **kw)
    for p, q in zip(ga[:-1], a[:-1]):
        bad = int(((p.double().cpu() - q).abs() > tol * float(q.abs().max())).sum())
        assert bad <= kink_budget and rel(p, q) < (100 if kink_budget else 1) * tol
    # and the reference-shaped discrete twin
    opt = J.optim.sgd(dt)
    b, st = devs(acts), opt.init(devs(acts))
    for _ in range(T):
        r = J.update_bpc_activities(gtd, gbu, b, opt, st, dev(y), input=gx, skip_model=gsk, **kw)
        assert set(r) == {"energy", "activities", "grads", "opt_state"}
        b, st = r["activities"], r["opt_state"]
    for p, q in zip(b[:-1], ga[:-1]):
        assert rel(p, q) < (1e-6 if tol < 1e-4 else tol)  # (bf16: the twin re-rounds its operand copies every call)


@pytest.mark.parametrize("act", ["relu", "tanh", "gelu"])
def test_bpc_small(act):
    c = _bpc_case(6, 12, 4, 9, act, 5, use_bias=True)
    _check_bpc(*c, TOL32)
    _check_bpc(*c, TOL32, af=1e-3, ab=0.7)
    _check_bpc_infer(*c, TOL32, T=6, dt=0.4 * 5, af=1e-3)


def test_bpc_skips_ragged():
    c = _bpc_case(7, 13, 5, 3, "tanh", 3, skips=True)
    _check_bpc(*c, TOL32)
    _check_bpc_infer(*c, TOL32, T=4, dt=0.9)


def test_bpc_two_layer_minimum():
    _check_bpc(*_bpc_case(5, 8, 2, 4, "silu", 4), TOL32)


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f32x3"])
def test_bpc_fused_train_step_equals_solve_then_param_grads(dtype):
    """jpcb_bpc_train_step (T Euler steps + both models' gradients + energy pairs, one call) against the two calls it
    fuses, and -- on the fp32 paths -- against the oracle's loop (examples/bidirectional_pc.ipynb cell 8)."""
    import jpc_b200 as J
    J.set_compute_dtype(dtype)
    tol = TOLBF if dtype == "bf16" else TOL32
    td, bu, sk, x, y, acts = _bpc_case(16, 64, 4, 24, "tanh", 32, use_bias=True) if dtype != "f32" else _bpc_case(6, 12, 4, 9, "tanh", 5, use_bias=True)
    gtd, gbu = to_gpu_model(td), to_gpu_model(bu)
    gx, gy, T, dt = dev(x), dev(y), 5, 0.6
    A1 = J.solve_bpc_inference(gtd, gbu, devs(acts), gy, input=gx, n_steps=T, dt=dt)
    td1, bu1 = J.compute_bpc_param_grads(gtd, gbu, A1, gy, x=gx)
    E1 = J.bpc_energy_fn(gtd, gbu, A1, gy, x=gx, record_layers=True)
    A2, td2, bu2, E2 = J.bpc_train_step(gtd, gbu, devs(acts), gy, input=gx, n_steps=T, dt=dt, record_energies=True)
    same = 1e-6 if dtype != "bf16" else 1e-2   # (bf16: the two-call form re-rounds the activities' operand copies)
    for p, q in zip(A2[:-1], A1[:-1]):
        assert rel(p, q.double().cpu()) < 1e-6
    assert rel(E2, E1.double().cpu()) < same
    for got, want in ((td2, td1), (bu2, bu1)):
        for l in range(len(want)):
            assert rel(got[l][1].weight, want[l][1].weight.double().cpu()) < same
            assert rel(got[l][1].bias, want[l][1].bias.double().cpu()) < same
    a = [t.clone() for t in acts]
    for _ in range(T):
        _, g = O.compute_bpc_activity_grad(td, bu, a, y, x=x, skips=sk)
        a = [p - dt * q for p, q in zip(a, g)]
    tdg, bug = O.compute_bpc_param_grads(td, bu, a, y, x=x, skips=sk)
    for got, want in ((td2, tdg), (bu2, bug)):
        for l, q in enumerate(want):
            assert rel(got[l][1].weight, q["W"]) < tol
            assert rel(got[l][1].bias, q["b"]) < tol
    assert rel(E2, O.bpc_energy_fn(td, bu, a, y, x=x, skips=sk, record_layers=True)) < tol


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f32x3"])
def test_bpc_unclamped_input(dtype):
    """x = None (jpc/_core/_energies.py:324): activities[0] is the input of top_down_model[0] AND the target of
    bottom_up_model[0], so its gradient has two more terms than in the clamped case and the operand copy phi_0(z_0) moves with
    every Euler step.  Energies, activity gradients, both models' parameter gradients, T fused steps and the fused train step
    against the oracle."""
    import jpc_b200 as J
    J.set_compute_dtype(dtype)
    tol = TOLBF if dtype == "bf16" else TOL32
    if dtype == "f32":
        cases = [_bpc_case(12, 12, 4, 9, "tanh", 5, use_bias=True), _bpc_case(13, 13, 5, 3, "silu", 3, skips=True)]
    else:
        cases = [_bpc_case(64, 64, 4, 24, "tanh", 32, use_bias=True), _bpc_case(72, 72, 5, 20, "tanh", 136, skips=True)]
    cases[0][0][0]["act"] = "tanh"   # (make_mlp leaves the first layer's input linear: one case with a real phi_0')
    for td, bu, sk, x, y, acts in cases:
        B = y.shape[0]
        _check_bpc(td, bu, sk, None, y, acts, tol)
        _check_bpc(td, bu, sk, None, y, acts, tol, af=0.5, ab=0.7)
        _check_bpc_infer(td, bu, sk, None, y, acts, tol, T=5, dt=0.4)
        _check_bpc_infer(td, bu, sk, None, y, acts, tol, T=3, dt=0.05 * B, af=0.3)
        # the fused train step: T steps + gradients + energies at the solution
        T, dt = 4, 0.3
        a = [t.clone() for t in acts]
        for _ in range(T):
            _, g = O.compute_bpc_activity_grad(td, bu, a, y, x=None, skips=sk)
            a = [p - dt * q for p, q in zip(a, g)]
        tdg, bug = O.compute_bpc_param_grads(td, bu, a, y, x=None, skips=sk)
        gtd, gbu, gsk = to_gpu_model(td), to_gpu_model(bu), _gpu_skip(sk)
        A2, td2, bu2, E2 = J.bpc_train_step(gtd, gbu, devs(acts), dev(y), input=None, n_steps=T, dt=dt, skip_model=gsk, record_energies=True)
        gate_all(A2[:-1], a[:-1], tol, "bpc_free.train_step.activities")
        for got, want in ((td2, tdg), (bu2, bug)):
            for l, q in enumerate(want):
                assert rel(got[l][1].weight, q["W"]) < tol
                if q["b"] is not None:
                    assert rel(got[l][1].bias, q["b"]) < tol
        assert rel(E2, O.bpc_energy_fn(td, bu, a, y, x=None, skips=sk, record_layers=True)) < tol
    # limits: activities[0] must fit top_down_model[0]'s input
    td, bu, sk, x, y, acts = _bpc_case(6, 12, 3, 9, "tanh", 4)
    with pytest.raises(ValueError):
        J.bpc_energy_fn(to_gpu_model(td), to_gpu_model(bu), devs(acts), dev(y), x=None)


def test_bpc_update_params_keys():
    import jpc_b200 as J
    td, bu, sk, x, y, acts = _bpc_case(6, 12, 3, 9, "relu", 4)
    gtd, gbu = to_gpu_model(td), to_gpu_model(bu)
    o1, o2 = J.optim.adam(1e-3), J.optim.adam(1e-3)
    r = J.update_bpc_params(gtd, gbu, devs(acts), o1, o2, o1.init(gtd), o2.init(gbu), dev(y), input=dev(x))
    assert set(r) == {"models", "grads", "opt_states"}
    assert len(r["models"]) == 2 and len(r["grads"]) == 2 and len(r["opt_states"]) == 2


def test_config5_bpc_full_size():
    """BASELINE config 5: bPC width 1024 depth 6 (10 -> 784), batch 1024, 30 Euler steps, fp32."""
    c = _bpc_case(10, 1024, 6, 784, "relu", 1024)
    _check_bpc(*c, TOL32)
    _check_bpc_infer(*c, TOL32, T=30, dt=0.5)


# ---- cross-entropy output energy (jpc/_core/_energies.py:107-109) ------------------------------
def _check_ce(om, sk, x, y, tol, param_type="sp", lam=None, fused=("euler", 3.0, 0.5)):
    import jpc_b200 as J
    gm, gsk = to_gpu_model(om), _gpu_skip(sk)
    gx, gy = dev(x), dev(y)
    a0 = O.init_activities_with_ffwd(om, x, skips=sk, param_type=param_type)
    acts = [a.float().double() for a in perturb(a0)]
    gacts = devs(acts)
    okw = dict(x=x, loss="ce", param_type=param_type, output_energy_scaling=lam)
    E = O.pc_energy_fn(om, sk, acts, y, record_layers=True, **okw)
    gE = J.pc_energy_fn((gm, gsk), gacts, gy, x=gx, loss="ce", param_type=param_type, output_energy_scaling=lam, record_layers=True)
    assert rel(gE, E) < tol
    F, g = O.compute_pc_activity_grad(om, sk, acts, y, **okw)
    gF, gg = J.compute_pc_activity_grad((gm, gsk), gacts, gy, x=gx, loss_id="ce", param_type=param_type, output_energy_scaling=lam)
    assert abs(float(gF) - float(F)) <= tol * abs(float(F))
    gate_all(gg, g, tol, "check_ce.activity_grads")
    pg = O.compute_pc_param_grads(om, sk, acts, y, **okw)
    gpg, _ = J.compute_pc_param_grads((gm, gsk), gacts, gy, x=gx, loss_id="ce", param_type=param_type, output_energy_scaling=lam)
    for l, q in enumerate(pg):
        assert rel(gpg[l][1].weight, q["W"]) < tol
        if q["b"] is not None:
            assert rel(gpg[l][1].bias, q["b"]) < tol
    # fused train step: loss is cross_entropy_loss at the init (jpc/_train.py:159-164)
    solver, max_t1, dt = fused
    ref = O.pc_train_step(om, sk, x, y, solver=solver, max_t1=max_t1, dt=dt, loss_id="ce", param_type=param_type)
    opt = J.optim.sgd(0.0)
    out = J.make_pc_step(gm, opt, opt.init((gm, gsk)), gy, input=gx, loss_id="ce", param_type=param_type,
                         ode_solver=J.Euler() if solver == "euler" else J.Heun(), max_t1=max_t1, dt=dt,
                         stepsize_controller=J.ConstantStepSize(), skip_model=gsk)
    assert abs(float(out["loss"]) - float(ref["loss"])) <= tol * abs(float(ref["loss"]))
    assert rel(out["energies"], ref["energies"]) < tol
    gate_all([p[0] for p in out["activities"][:-1]], ref["activities"][:-1], tol, "check_ce.activities")


def test_cross_entropy_fp32():
    om, sk, x, y = _case(10, 20, 3, 5, "tanh", 6, use_bias=True)            # one-hot targets
    _check_ce(om, sk, x, y, TOL32)
    _check_ce(om, sk, x, y, TOL32, lam=2.5, fused=("heun", 2.0, 0.5))
    om, sk, x, y = _case(12, 16, 4, 37, "relu", 5, onehot_out=False)         # targets that do not sum to 1
    _check_ce(om, sk, x, y.abs(), TOL32)
    om, sk, x, y = _case(784, 300, 3, 10, "tanh", 64)                         # config-1 shape with the ce energy
    _check_ce(om, sk, x, y, TOL32, fused=("euler", 10.0, 0.5))


def test_cross_entropy_bf16():
    import jpc_b200 as J
    J.set_compute_dtype("bf16")
    om, sk, x, y = _case(256, 512, 3, 16, "tanh", 256, use_bias=True)
    _check_ce(om, sk, x, y, TOLBF, fused=("euler", 2.0, 0.5))


# ---- adaptive solve: Heun + PIDController, the reference's default ----------------
def _adaptive_case(om, sk, x, y, tol, max_t1, dt0, rtol=1e-3, atol=1e-3, param_type="sp", ad=0.0):
    import jpc_b200 as J
    gm, gsk = to_gpu_model(om), _gpu_skip(sk)
    gx, gy = dev(x), dev(y)
    a0 = O.init_activities_with_ffwd(om, x, skips=sk, param_type=param_type)
    ref, st = O.solve_inference_adaptive(om, sk, a0, y, x=x, max_t1=max_t1, dt0=dt0, rtol=rtol, atol=atol, param_type=param_type,
                                         activity_decay=ad)
    sol = J.solve_inference((gm, gsk), devs(a0), gy, input=gx, param_type=param_type, solver=J.Heun(), max_t1=max_t1, dt=dt0,
                            stepsize_controller=J.PIDController(rtol=rtol, atol=atol), activity_decay=ad)
    got = J.solve_inference.last_stats
    assert (got["accepted"], got["rejected"], got["event"]) == (st["accepted"], st["rejected"], st["event"])
    for p in sol:
        assert p.ndim == 3 and p.shape[0] == 1
    gate_all([p[0] for p in sol[:-1]], ref[:-1], tol, "adaptive.activities")
    assert float((sol[-1][0].double().cpu() - ref[-1]).abs().max()) <= tol * max(1e-30, max(float(t.abs().max()) for t in ref))
    return st


def test_adaptive_heun_pid_matches_oracle_controller():
    om, sk, x, y = _case(10, 20, 3, 5, "tanh", 8, use_bias=True)
    st = _adaptive_case(om, sk, x, y, 2e-5, 5.0, None)            # Hairer initial step
    assert st["accepted"] >= 3
    _adaptive_case(om, sk, x, y, 2e-5, 5.0, 0.05)                  # given dt0
    st = _adaptive_case(om, sk, x, y, 2e-5, 3.0, 50.0, rtol=1e-5, atol=1e-5)   # first step rejected
    assert st["rejected"] >= 1
    om, sk, x, y = _case(12, 16, 6, 3, "relu", 5, param_type="mupc", skips=True)
    _adaptive_case(om, sk, x, y, 2e-5, 4.0, None, param_type="mupc")
    om, sk, x, y = _case(10, 20, 3, 5, "tanh", 8, use_bias=True)
    _adaptive_case(om, sk, x, y, 2e-5, 5.0, None, ad=0.5)          # activity decay: the dummy output slot decays too


def test_make_pc_step_default_arguments_run_the_adaptive_solve():
    """make_pc_step(model, optim, opt_state, output, input=...) with the reference's defaults (Heun, PIDController
    1e-3, max_t1=20, dt=None) -- the call every example notebook makes."""
    import jpc_b200 as J
    om, sk, x, y = _case(784, 300, 3, 10, "relu", 64, use_bias=True)
    gm = to_gpu_model(om)
    gx, gy = dev(x), dev(y)
    opt = J.optim.adam(1e-3)
    out = J.make_pc_step(gm, opt, opt.init((gm, None)), gy, input=gx)
    a0 = O.init_activities_with_ffwd(om, x)
    ref, st = O.solve_inference_adaptive(om, None, a0, y, x=x, max_t1=20.0)
    assert abs(float(out["loss"]) - float(O.mse_loss(a0[-1], y))) <= TOL32 * float(O.mse_loss(a0[-1], y))
    E = O.pc_energy_fn(om, None, ref, y, x=x, record_layers=True)
    assert rel(out["energies"], E) < 1e-4
    gate_all([p[0] for p in out["activities"][:-1]], ref[:-1], 2e-5, "default_args.activities")
    pg = O.compute_pc_param_grads(om, None, ref, y, x=x)
    g, _ = J.compute_pc_param_grads((gm, None), [a[0].contiguous() for a in out["activities"]], gy, x=gx)
    for l, q in enumerate(pg):
        assert rel(g[l][1].weight, q["W"]) < 1e-4
    with pytest.raises(RuntimeError):  # Euler has no error estimate: diffrax refuses, so do we
        J.make_pc_step(gm, opt, opt.init((gm, None)), gy, input=gx, ode_solver=J.Euler())


def test_bpc_bf16_tensor_core_path():
    """bPC on the tcgen05 kernels (bf16 operands, 2e-2): feature dims that are not multiples of 8 (10, 28) are
    zero-padded by the host mirror and sliced back."""
    import jpc_b200 as J
    J.set_compute_dtype("bf16")
    c = _bpc_case(10, 256, 4, 28, "tanh", 128, use_bias=True)
    _check_bpc(*c, TOLBF)
    _check_bpc(*c, TOLBF, af=0.5, ab=0.7)
    _check_bpc_infer(*c, TOLBF, T=6, dt=0.1 * 128)  # SURVEY 8d stress step 0.1 * B
    c = _bpc_case(16, 128, 5, 72, "relu", 136, skips=True)
    _check_bpc(*c, TOLBF)
    _check_bpc_infer(*c, TOLBF, T=4, dt=0.5)          # (relu' flips under bf16 rounding near z = 0: reference-sized steps)
    c = _bpc_case(16, 128, 5, 72, "tanh", 136, skips=True)
    _check_bpc_infer(*c, TOLBF, T=4, dt=0.1 * 136)    # stress step with a smooth activation


def test_config5_bpc_full_size_bf16():
    import jpc_b200 as J
    J.set_compute_dtype("bf16")
    c = _bpc_case(10, 1024, 6, 784, "relu", 1024)
    _check_bpc(*c, TOLBF)
    _check_bpc_infer(*c, TOLBF, T=30, dt=0.5)


# ---- tensor-core GEMM shape sweep (tile / TMA box boundaries) -----------------------------------
@pytest.mark.parametrize("B,d_in,d_out", [(8, 8, 8), (8, 64, 8), (136, 72, 200), (128, 64, 256), (256, 2048, 264),
                                          (264, 520, 136), (520, 8, 1032), (1024, 136, 24), (40, 1000, 512)])
def test_tc_gemm_shape_sweep(B, d_in, d_out):
    """One FFWD GEMM + bias through the tcgen05 path for shapes around the 256x256x64 tile and the TMA boxes
    (M, N, K below / at / above one tile, K not a multiple of 64), against fp64 on the bf16-rounded operands."""
    import jpc_b200 as J
    from jpc_b200 import nn
    J.set_compute_dtype("bf16")
    g = torch.Generator().manual_seed(B * 7 + d_in * 3 + d_out)
    W = (torch.rand(d_out, d_in, generator=g, dtype=torch.float64) * 2 - 1) / math.sqrt(d_in)
    b = torch.rand(d_out, generator=g, dtype=torch.float64) - 0.5
    x = torch.randn(B, d_in, generator=g, dtype=torch.float64)
    W2 = (torch.rand(d_out, d_out, generator=g, dtype=torch.float64) * 2 - 1) / math.sqrt(d_out)
    layers = []
    for Wl, bl, act in ((W, b, "linear"), (W2, None, "tanh")):
        lin = nn.Linear.__new__(nn.Linear)
        lin.weight, lin.bias = dev(Wl), (None if bl is None else dev(bl))
        lin.out_features, lin.in_features = Wl.shape
        lin.use_bias = bl is not None
        layers.append(nn.Sequential([nn.Lambda(nn.get_act_fn(act)), lin]))
    # mixed bias is not on the path: give the second layer a zero bias
    layers[1].layers[1].bias = torch.zeros(d_out, device="cuda")
    layers[1].layers[1].use_bias = True
    out = J.init_activities_with_ffwd(layers, dev(x))
    r = lambda t: t.float().bfloat16().double()
    want = r(x) @ r(W).T + b.float().double()
    assert float((out[0].double().cpu() - want).abs().max()) < 2e-5 * max(float(want.abs().max()), 1.0)
    want2 = r(torch.tanh(want)) @ r(W2).T
    assert float((out[1].double().cpu() - want2).abs().max()) < 2e-2 * max(float(want2.abs().max()), 1.0)


def test_bf16_path_odd_feature_dims_are_zero_padded():
    """MNIST-shaped net on the tensor-core path: 10 classes / a 20-wide layer are not multiples of 8; the host
    mirror zero-pads feature dims and slices every result back (energies, gradients, activities unchanged)."""
    import jpc_b200 as J
    J.set_compute_dtype("bf16")
    om, sk, x, y = _case(784, 260, 4, 10, "tanh", 128, use_bias=True)
    _check_all(om, sk, x, y, TOLBF)
    _check_fused(om, sk, x, y, "euler", 5.0, 0.5, TOLBF)
    _check_fused(om, sk, x, y, "heun", 2.0, 0.5, TOLBF)
    om, sk, x, y = _case(20, 36, 5, 12, "tanh", 128, skips=True)
    _check_all(om, sk, x, y, TOLBF)
    with pytest.raises(NotImplementedError):  # zero-padded logits would enter the softmax
        J.pc_energy_fn((to_gpu_model(om), None), devs(O.init_activities_with_ffwd(om, x)), dev(y), x=dev(x), loss="ce")


def test_training_loop_learns_a_synthetic_task():
    """End to end through the public API, the way the reference's discriminative_pc example trains: repeated
    make_pc_step (default Heun + PID solve on step 0, then fixed Euler steps) with adam lowers the test loss of a
    teacher-labelled synthetic classification task and lifts accuracy far above chance."""
    import jpc_b200 as J
    J.set_compute_dtype("f32")
    g = torch.Generator().manual_seed(0)
    d_in, d_out, B = 32, 4, 256
    teacher = torch.randn(d_in, d_out, generator=g)
    def batch():
        x = torch.randn(B, d_in, generator=g)
        y = torch.nn.functional.one_hot((x @ teacher).argmax(1), d_out).float()
        return x.cuda(), y.cuda()
    model = J.make_mlp(1, d_in, 64, 3, d_out, "relu", use_bias=True, device="cuda")
    opt = J.optim.adam(3e-3)
    st = opt.init((model, None))
    xt, yt = batch()
    loss0, acc0 = J.test_discriminative_pc(model, yt, xt)
    for it in range(150):
        x, y = batch()
        kw = {} if it == 0 else dict(ode_solver=J.Euler(), stepsize_controller=J.ConstantStepSize(), max_t1=5.0, dt=0.5)
        out = J.make_pc_step(model, opt, st, y, input=x, **kw)
        model, st = out["model"], out["opt_state"]
        assert torch.isfinite(out["loss"])
    loss1, acc1 = J.test_discriminative_pc(model, yt, xt)
    assert float(loss1) < 0.6 * float(loss0)
    assert float(acc1) > 70.0 > 40.0 > float(acc0) - 15.0


# ---- recorded trajectories: record_iters / record_every / record_energies -----------
@pytest.mark.parametrize("solver", ["euler", "heun"])
def test_solve_inference_record_iters_and_record_every(solver):
    import jpc_b200 as J
    om, sk, x, y = _case(10, 20, 3, 5, "tanh", 8, use_bias=True)
    gm, gx, gy = to_gpu_model(om), dev(x), dev(y)
    a0 = [a.float().double() for a in perturb(O.init_activities_with_ffwd(om, x))]
    traj = []
    O.solve_inference(om, sk, a0, y, x=x, solver=solver, max_t1=2.0, dt=0.3, use_closed_form=True, event_atol=0.0,
                      event_rtol=0.0, trajectory=traj)
    kw = dict(input=gx, solver=J.Euler() if solver == "euler" else J.Heun(), max_t1=2.0, dt=0.3,
              stepsize_controller=J.ConstantStepSize())
    # SaveAt(steps=True): (4096, B, d), one row per step (7 here, the last one clipped), inf after
    sol = J.solve_inference((gm, None), devs(a0), gy, record_iters=True, **kw)
    assert int(J.get_t_max(sol)) == 6
    for l, s in enumerate(sol):
        assert tuple(s.shape) == (4096,) + tuple(a0[l].shape)
        assert bool(torch.isinf(s[7:]).all())
        for i in range(7):
            assert rel(s[i], traj[i + 1][1][l]) < TOL32
    # SaveAt(ts=arange(0, max_t1, 1), t1=True): rows at t = 0, 1 (interpolated inside a step), then the final state
    sol = J.solve_inference((gm, None), devs(a0), gy, record_every=1, **kw)
    want = O.saveat_ts(traj, [0.0, 1.0])
    for l, s in enumerate(sol):
        assert tuple(s.shape) == (3,) + tuple(a0[l].shape)
        assert rel(s[0], a0[l]) < 1e-7 and rel(s[1], want[1][l]) < TOL32 and rel(s[2], traj[-1][1][l]) < TOL32


def test_make_pc_step_record_energies_and_adaptive_recording():
    import jpc_b200 as J
    om, sk, x, y = _case(12, 24, 3, 6, "relu", 8, use_bias=True)
    gm, gx, gy = to_gpu_model(om), dev(x), dev(y)
    opt = J.optim.sgd(1e-2)
    out = J.make_pc_step(gm, opt, opt.init((gm, None)), gy, input=gx, ode_solver=J.Euler(), max_t1=3, dt=0.5,
                         stepsize_controller=J.ConstantStepSize(), record_energies=True, activity_norms=True)
    a0 = O.init_activities_with_ffwd(om, x)
    traj = []
    O.solve_inference(om, sk, a0, y, x=x, solver="euler", max_t1=3.0, dt=0.5, use_closed_form=True, event_atol=0.0,
                      event_rtol=0.0, trajectory=traj)
    assert int(out["t_max"]) == 5 and tuple(out["energies"].shape) == (3, 500)
    for t in range(5):  # jpc/_utils.py:203-216: steps 0..t_max-1, layer rows reversed
        E = O.pc_energy_fn(om, None, traj[t + 1][1], y, x=x, record_layers=True)
        assert rel(out["energies"][:, t], torch.flip(E, [0])) < 1e-5
    assert float(out["energies"][:, 5:].abs().max()) == 0.0
    # the parameter update uses the last recorded row
    fin = traj[-1][1]
    pg = O.compute_pc_param_grads(om, None, fin, y, x=x)
    new = out["model"]
    for l, q in enumerate(pg):
        assert rel(new[l][1].weight, om[l]["W"] - 1e-2 * q["W"]) < 1e-6
    norms = torch.stack([a.norm(dim=-1).mean() for a in fin])
    assert rel(out["activity_norms"], norms) < 1e-5
    # default adaptive solve with every accepted step recorded
    out = J.make_pc_step(gm, opt, opt.init((gm, None)), gy, input=gx, max_t1=5, record_activities=True)
    ref, st = O.solve_inference_adaptive(om, None, a0, y, x=x, max_t1=5.0)
    tm = int(out["t_max"])
    assert tm == st["accepted"] - 1
    for p, q in zip(out["activities"][:-1], ref[:-1]):
        assert tuple(p.shape)[0] == 4096 and rel(p[tm], q) < 1e-4


# ---- fused parameter update ---------------------------------------------------------
@pytest.mark.parametrize("name", ["sgd", "adam"])
def test_fused_optimiser_matches_optax_protocol(name):
    """jpcb_optim_sgd / jpcb_optim_adam (one pass) against the optax-protocol path (update + apply_updates) and
    the fp64 oracle, over ragged tensor sizes, an unaligned view, an empty tensor and > 32 tensors (two launches)."""
    import jpc_b200 as J
    from jpc_b200.optim import apply_updates, update_and_apply
    g_ = torch.Generator().manual_seed(0)
    sizes = [(7, 5), (1,), (0,), (33, 129), (1024, 64)] + [(3, k + 1) for k in range(36)]
    base = torch.randn(8, generator=g_, dtype=torch.float64).float().cuda()
    p64 = [torch.randn(s, generator=g_, dtype=torch.float64) for s in sizes] + [base[1:8].double().cpu()]
    params = [dev(p) for p in p64[:-1]] + [base[1:8]]  # last leaf: contiguous but not 16-byte aligned
    opt = J.optim.sgd(0.3) if name == "sgd" else J.optim.adam(1e-2)
    st_f, st_p, st_o = opt.init(params), opt.init(params), None
    pf, pp, po = params, params, p64
    n0 = J._lib.lib.jpcb_launch_count()
    for it in range(3):
        g64 = [torch.randn(p.shape, generator=g_, dtype=torch.float64) for p in p64]
        g = devs(g64)
        pf, st_f = update_and_apply(opt, g, st_f, pf)
        u, st_p = opt.update(g, st_p, pp)
        pp = apply_updates(pp, u)
        if name == "sgd":
            po = O.sgd_update(po, g64, 0.3)
        else:
            po, st_o = O.adam_update(po, g64, st_o, 1e-2)
    assert J._lib.lib.jpcb_launch_count() - n0 == 6  # 42 tensors -> two launches per update: the fused path ran
    for a, b, c in zip(pf, pp, po):
        if a.numel():
            assert rel(a, c) < 2e-6 and rel(b, c) < 2e-6
    assert params[0] is not pf[0] and rel(params[0], p64[0]) < 1e-7  # functional: inputs untouched
    if name == "adam":
        for a, c in zip(st_f["mu"] + st_f["nu"], st_o["mu"] + st_o["nu"]):
            if a.numel():
                assert rel(a, c) < 2e-6
    # CPU tensors: the protocol path (no device kernel) still serves them
    cp = [torch.ones(3)]
    out, _ = update_and_apply(J.optim.sgd(0.5), [torch.ones(3)], (), cp)
    assert torch.equal(out[0], torch.full((3,), 0.5))


# ---- unsupervised PC: input=None, z_0 is a free activity -----------------------------
def _unsup_case(dims_in, width, depth, d_out, act, B, use_bias, act0="linear", seed=0):
    om, sk, x, y = _case(dims_in, width, depth, d_out, act, B, use_bias=use_bias, seed=seed)
    om[0]["act"] = act0
    g = torch.Generator().manual_seed(5)
    dims = [dims_in] + [l["W"].shape[0] for l in om]
    acts = [(0.5 * torch.randn(B, w, generator=g, dtype=torch.float64)).float().double() for w in dims]
    return om, y, acts


def _check_unsup(om, y, acts, tol, ad=0.0, lam=None, solves=True):
    import jpc_b200 as J
    gm, gy, ga = to_gpu_model(om), dev(y), devs(acts)
    fm = f32_model(om)
    kw = dict(x=None, activity_decay=ad, output_energy_scaling=lam)
    E = O.pc_energy_fn(fm, None, acts, y, record_layers=True, **kw)
    assert rel(J.pc_energy_fn((gm, None), ga, gy, record_layers=True, **kw), E) < tol
    F, g = O.compute_pc_activity_grad(fm, None, acts, y, **kw)
    gF, gg = J.compute_pc_activity_grad((gm, None), ga, gy, **kw)
    assert abs(float(gF) - float(F)) <= tol * abs(float(F))
    assert len(gg) == len(om) + 1
    gate_all(gg, g, tol, "check_unsup.activity_grads")
    pg = O.compute_pc_param_grads(fm, None, acts, y, **kw)
    gp, _ = J.compute_pc_param_grads((gm, None), ga, gy, **kw)
    for l, q in enumerate(pg):
        assert rel(gp[l][1].weight, q["W"]) < tol
        if q["b"] is not None:
            assert rel(gp[l][1].bias, q["b"]) < tol
    if not solves:
        return
    for solver in ("euler", "heun"):
        ref, _ = O.solve_inference(fm, None, acts, y, x=None, solver=solver, max_t1=2.0, dt=0.3 * y.shape[0], event_atol=0.0,
                                   event_rtol=0.0, activity_decay=ad, output_energy_scaling=lam)
        sol = J.solve_inference((gm, None), ga, gy, input=None, solver=J.Euler() if solver == "euler" else J.Heun(),
                                max_t1=2.0, dt=0.3 * y.shape[0], stepsize_controller=J.ConstantStepSize(), activity_decay=ad,
                                output_energy_scaling=lam)
        assert len(sol) == len(om) + 1
        gate_all([p[0] for p in sol], ref, 4 * tol, "check_unsup.solve")
        assert float((sol[0][0].double().cpu() - acts[0]).abs().max()) > 1e-3  # z_0 did move


def test_unsupervised_free_input_fp32():
    om, y, acts = _unsup_case(12, 20, 3, 6, "tanh", 8, True)
    _check_unsup(om, y, acts, TOL32)
    _check_unsup(om, y, acts, TOL32, ad=0.05, lam=0.7)
    om, y, acts = _unsup_case(9, 17, 4, 5, "relu", 6, False, act0="tanh")   # ragged dims, non-linear first layer
    _check_unsup(om, y, acts, TOL32)
    import jpc_b200 as J
    gm, gy, ga = to_gpu_model(om), dev(y), devs(acts)
    with pytest.raises(ValueError):   # L entries instead of L + 1
        J.pc_energy_fn((gm, None), ga[1:], gy, x=None)
    with pytest.raises(ValueError):   # the reference's scalings need the input's width
        J.pc_energy_fn((gm, None), ga, gy, x=None, param_type="mupc")


def test_unsupervised_free_input_bf16():
    import jpc_b200 as J
    om, y, acts = _unsup_case(64, 128, 3, 32, "tanh", 64, True, act0="relu")
    J.set_compute_dtype("bf16")
    _check_unsup(om, y, acts, TOLBF)
    om, y, acts = _unsup_case(50, 100, 3, 10, "relu", 32, False)           # padded odd dims, linear first layer
    _check_unsup(om, y, acts, TOLBF)


def test_unsupervised_adaptive_and_make_pc_step():
    import jpc_b200 as J
    om, y, acts = _unsup_case(12, 24, 3, 6, "tanh", 8, True)
    gm, gy, ga = to_gpu_model(om), dev(y), devs(acts)
    fm = f32_model(om)
    # default solve (Heun + PID) from given activities
    ref, st = O.solve_inference_adaptive(fm, None, acts, y, x=None, max_t1=6.0)
    sol = J.solve_inference((gm, None), ga, gy, input=None, max_t1=6)
    got = J.solve_inference.last_stats
    assert (got["accepted"], got["rejected"]) == (st["accepted"], st["rejected"])
    gate_all([p[0] for p in sol], ref, 2e-5, "unsup_adaptive.activities")
    # make_pc_step(input=None): random-normal init (same seed -> same draw), no loss, L + 1 activities
    sizes = [12, 24, 24, 6]
    a0 = J.init_activities_from_normal(3, sizes, "unsupervised", 8, sigma=0.05)
    assert [tuple(a.shape) for a in a0] == [(8, w) for w in sizes]
    assert 0.03 < float(torch.cat([a.flatten() for a in a0]).std()) < 0.07
    opt = J.optim.sgd(1e-2)
    out = J.make_pc_step(gm, opt, opt.init((gm, None)), gy, key=3, layer_sizes=sizes, batch_size=8, ode_solver=J.Euler(),
                         max_t1=3, dt=0.5, stepsize_controller=J.ConstantStepSize())
    assert out["loss"] is None and len(out["activities"]) == 4
    a064 = [a.double().cpu() for a in a0]
    ref, _ = O.solve_inference(fm, None, a064, y, x=None, solver="euler", max_t1=3.0, dt=0.5, event_atol=0.0, event_rtol=0.0)
    for p, q in zip(out["activities"], ref):
        assert rel(p[0], q) < 1e-5
    assert rel(out["energies"], O.pc_energy_fn(fm, None, ref, y, x=None, record_layers=True)) < 1e-5
    pg = O.compute_pc_param_grads(fm, None, ref, y, x=None)
    for l, q in enumerate(pg):
        assert rel(out["model"][l][1].weight, fm[l]["W"] - 1e-2 * q["W"]) < 1e-6
    with pytest.raises(ValueError):   # jpc/_train.py:136-141
        J.make_pc_step(gm, opt, opt.init((gm, None)), gy)


# ---- hybrid PC: amortiser + generator -------------------------------------------------
def _hpc_models(d_in, width, depth, d_out, act, use_bias, seed=0):
    """generator d_in -> d_out and amortiser d_out -> d_in of the same depth (examples/hybrid_pc.ipynb)."""
    gen = O.make_mlp(seed, d_in, width, depth, d_out, act, use_bias=use_bias)
    amo = O.make_mlp(seed + 1, d_out, width, depth, d_in, act, use_bias=use_bias)
    return gen, amo


@pytest.mark.parametrize("supervised", [True, False])
def test_hybrid_pc_step_matches_oracle(supervised):
    import jpc_b200 as J
    B = 8
    gen, amo = _hpc_models(10, 24, 3, 14, "tanh", True)
    x, y = data(B, 10, 14, onehot_out=False, onehot_in=True)
    ggen, gamo = to_gpu_model(gen), to_gpu_model(amo)
    gx, gy = dev(x), dev(y)
    fgen, famo = f32_model(gen), f32_model(amo)
    inp, ginp = (x, gx) if supervised else (None, None)
    # amortised initialisation
    aa = O.init_activities_with_amort(famo, fgen, y)
    gaa = J.init_activities_with_amort(gamo, ggen, gy)
    assert len(gaa) == len(aa) == 4
    for p, q in zip(gaa, aa):
        assert rel(p, q) < TOL32
    # one hybrid step with a fixed-step solve
    opts = (J.optim.sgd(1e-2), J.optim.sgd(1e-2))
    states = (opts[0].init((ggen, None)), opts[1].init(gamo))
    out = J.make_hpc_step(ggen, gamo, opts, states, gy, input=ginp, ode_solver=J.Euler(), max_t1=3, dt=0.5,
                          stepsize_controller=J.ConstantStepSize(), record_energies=True)
    a0 = aa[1:] if supervised else aa
    traj = []
    eq, _ = O.solve_inference(fgen, None, a0, y, x=inp, solver="euler", max_t1=3.0, dt=0.5, event_atol=0.0, event_rtol=0.0,
                              trajectory=traj)
    assert int(out["t_max"]) == 5
    for p, q in zip(out["activities"][1], eq):
        assert rel(p[5], q) < TOL32
    for_amort = eq[::-1][1:]
    E = O.hpc_energy_fn(famo, for_amort, aa, y, inp)
    assert abs(float(out["energies"][1]) - float(E)) <= TOL32 * float(E)
    El = O.hpc_energy_fn(famo, for_amort, aa, y, inp, record_layers=True)
    assert rel(J.hpc_energy_fn(gamo, devs(for_amort), gaa, gy, ginp, record_layers=True), El) < TOL32
    gg = O.compute_pc_param_grads(fgen, None, eq, y, x=inp)
    ag = O.compute_hpc_param_grads(famo, for_amort, aa, y, inp)
    for l in range(3):
        assert rel(out["generator"][l][1].weight, fgen[l]["W"] - 1e-2 * gg[l]["W"]) < 1e-6
        assert rel(out["amortiser"][l][1].weight, famo[l]["W"] - 1e-2 * ag[l]["W"]) < 1e-6
        assert rel(out["amortiser"][l][1].bias, famo[l]["b"] - 1e-2 * ag[l]["b"]) < 1e-6
    jg = J.compute_hpc_param_grads(gamo, devs(for_amort), gaa, gy, ginp)
    for l in range(3):
        assert rel(jg[l][1].weight, ag[l]["W"]) < TOL32 and rel(jg[l][1].bias, ag[l]["b"]) < TOL32
    if supervised:
        assert abs(float(out["losses"][0]) - float(torch.mean((y - O.init_activities_with_ffwd(fgen, x)[-1]) ** 2))) < 1e-5
        assert abs(float(out["losses"][1]) - float(torch.mean((x - aa[0]) ** 2))) < 1e-5
    else:
        assert out["losses"][0] is None
        assert abs(float(out["losses"][1]) - float(torch.mean((for_amort[-1] - aa[0]) ** 2))) < 1e-5
    # evaluation helper runs end to end (default adaptive solves, amortised and random inits)
    accs = J.test_hpc(ggen, gamo, gy, gx, 0, [10, 24, 24, 14], B, max_t1=5)
    assert len(accs) == 4 and tuple(accs[3].shape) == (B, 14) and all(0.0 <= float(a) <= 100.0 for a in accs[:3])
    acc, preds = J.test_generative_pc(ggen, gy, gx, 0, [10, 24, 24, 14], B, max_t1=5)
    assert 0.0 <= float(acc) <= 100.0 and rel(preds, O.init_activities_with_ffwd(fgen, x)[-1]) < TOL32


def test_hybrid_pc_bf16():
    """(targets 0.5 away from the predictions: with nearly-converged targets the regression error is a small
    difference of bf16-rounded numbers and its relative accuracy is that cancellation's, not the kernel's)"""
    import jpc_b200 as J
    B = 32
    gen, amo = _hpc_models(16, 64, 3, 24, "relu", False)
    x, y = data(B, 16, 24, onehot_out=False)
    fgen, famo = f32_model(gen), f32_model(amo)
    aa = O.init_activities_with_amort(famo, fgen, y)
    eq = [a + 0.5 * torch.randn(a.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64) for a in aa[1:]]
    for_amort = eq[::-1][1:]
    J.set_compute_dtype("bf16")
    gamo = to_gpu_model(amo)
    ag = O.compute_hpc_param_grads(famo, for_amort, aa, y, x)
    jg = J.compute_hpc_param_grads(gamo, devs(for_amort), devs(aa), dev(y), dev(x))
    for l in range(3):
        assert rel(jg[l][1].weight, ag[l]["W"]) < TOLBF
    E = O.hpc_energy_fn(famo, for_amort, aa, y, x, record_layers=True)
    assert rel(J.hpc_energy_fn(gamo, devs(for_amort), devs(aa), dev(y), dev(x), record_layers=True), E) < TOLBF


# ---- error-reparameterised PC ------------------------------------------------------
def _check_epc(om, sk, x, y, tol, param_type="sp", loss="mse", kink_frac=0.0):
    """`kink_frac`: fraction of error-gradient entries allowed outside `tol`.  In ePC the activities are OUTPUTS of the
    forward GEMMs, so on the bf16 path a relu pre-activation within ~1e-3 of zero can land on the other side of the
    kink and flip relu'(z) for that one entry (measured, scripts/epc_check.py: 1-5 entries of ~3000 per tensor, none
    with tanh, none in fp32).  PC proper does not have this: its activities are fp32 master state."""
    import jpc_b200 as J
    gm, gsk, gx, gy = to_gpu_model(om), _gpu_skip(sk), dev(x), dev(y)
    g = torch.Generator().manual_seed(11)
    dims = [l["W"].shape[0] for l in om]
    errs = [(0.3 * torch.randn(x.shape[0], w, generator=g, dtype=torch.float64)).float().double() for w in dims]
    kw = dict(x=x, param_type=param_type)
    F, ge = O.compute_epc_error_grad(om, sk, errs, y, loss=loss, **kw)
    gF, gge = J.compute_epc_error_grad((gm, gsk), devs(errs), gy, x=gx, loss_id=loss, param_type=param_type)
    assert abs(float(gF) - float(F)) <= tol * abs(float(F))
    assert abs(float(J.epc_energy_fn((gm, gsk), devs(errs), gy, x=gx, loss=loss, param_type=param_type)) - float(F)) <= tol * abs(float(F))
    gmax = max(float(t.abs().max()) for t in ge)
    for p, q in zip(gge, ge):
        bad = ((p.double().cpu() - q).abs() > tol * gmax).double().mean()
        assert float(bad) <= kink_frac
    pg = O.compute_epc_param_grads(om, sk, errs, y, loss=loss, **kw)
    mg, sg = J.compute_epc_param_grads((gm, gsk), devs(errs), gy, x=gx, loss_id=loss, param_type=param_type)
    wmax = max(float(q["W"].abs().max()) for q in pg)
    for l, q in enumerate(pg):
        dW = (mg[l][1].weight.double().cpu() - q["W"]).abs()
        lim = tol * max(wmax, float(q["W"].abs().max()))
        assert float((dW > lim).double().mean()) <= kink_frac and float(dW.max()) < (10 if kink_frac else 1) * lim
        if q["b"] is not None:
            assert rel(mg[l][1].bias, q["b"]) < tol or float((mg[l][1].bias.double().cpu() - q["b"]).abs().max()) < tol * wmax
    # one error update + one parameter update, the reference's loop body
    opt = J.optim.sgd(0.1)
    res = J.update_epc_errors((gm, gsk), devs(errs), opt, opt.init(devs(errs)), gy, input=gx, loss_id=loss, param_type=param_type)
    for p, e0, q in zip(res["errors"], errs, gge):   # (the update applied to the library's own gradient)
        assert float((p.double().cpu() - (e0 - 0.1 * q.double().cpu())).abs().max()) <= 1e-5 * max(1.0, float(e0.abs().max()))
    res = J.update_epc_params((gm, gsk), devs(errs), opt, opt.init((gm, gsk)), gy, input=gx, loss_id=loss, param_type=param_type)
    assert set(res) == {"model", "skip_model", "grads", "opt_state"} and len(res["model"]) == len(om)


def test_epc_fp32():
    om, sk, x, y = _case(10, 20, 3, 5, "tanh", 8, use_bias=True)
    _check_epc(om, sk, x, y, TOL32)
    om, sk, x, y = _case(9, 17, 4, 6, "relu", 6, onehot_out=True)
    _check_epc(om, sk, x, y, TOL32, loss="ce")
    om, sk, x, y = _case(12, 16, 6, 3, "relu", 5, param_type="mupc", skips=True)
    _check_epc(om, sk, x, y, TOL32, param_type="mupc")
    import jpc_b200 as J
    errs = J.init_epc_errors([10, 20, 20, 5], 8)
    assert [tuple(e.shape) for e in errs] == [(8, 20), (8, 20), (8, 5)] and float(sum(e.abs().sum() for e in errs)) == 0.0
    # x = None: errors[0] plays z_0, L + 1 entries; the reference's penalty covers the first L - 1 list entries (z_0 in,
    # eps_{L-1} out) and scales the first layer by the width of errors[0]
    for ptype, skips in (("sp", False), ("mupc", True)):
        om, sk, x, y = _case(12, 16, 5, 4, "tanh", 6, use_bias=True, param_type=ptype, skips=skips)
        g = torch.Generator().manual_seed(7)
        full = [(0.4 * torch.randn(6, w, generator=g, dtype=torch.float64)).float().double() for w in [12, 16, 16, 16, 16, 4]]
        F, ge = O.compute_epc_error_grad(om, sk, full, y, x=None, param_type=ptype)
        pg = O.compute_epc_param_grads(om, sk, full, y, x=None, param_type=ptype)
        P = (to_gpu_model(om), _gpu_skip(sk))
        gF, gge = J.compute_epc_error_grad(P, devs(full), dev(y), x=None, param_type=ptype)
        assert abs(float(gF) - float(F)) <= TOL32 * abs(float(F)) and len(gge) == 6
        gmax = max(float(t.abs().max()) for t in ge)
        for p, q in zip(gge, ge):
            assert float((p.double().cpu() - q).abs().max()) <= TOL32 * gmax
        mg, _ = J.compute_epc_param_grads(P, devs(full), dev(y), x=None, param_type=ptype)
        wmax = max(float(q["W"].abs().max()) for q in pg)
        for l, q in enumerate(pg):
            assert float((mg[l][1].weight.double().cpu() - q["W"]).abs().max()) <= TOL32 * wmax


def test_epc_bf16():
    import jpc_b200 as J
    om, sk, x, y = _case(64, 128, 4, 32, "tanh", 64, use_bias=True)
    J.set_compute_dtype("bf16")
    _check_epc(om, sk, x, y, TOLBF)
    om, sk, x, y = _case(50, 96, 4, 10, "relu", 32, skips=True, onehot_out=False)
    _check_epc(om, sk, x, y, TOLBF, kink_frac=5e-3)


# ---- BASELINE config 4 at full size (the bench line's workload) ------------------------------------
def _bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _c4_recompute(om, A, x, y, B, faithful):
    """Per-layer errors, energies (reference order) and weight gradients recomputed on the host from the activities the
    GPU returned.  faithful=False: fp64, exact operands (the reference's arithmetic).  faithful=True: the GPU path's
    operand rounding (bf16 act(z), bf16 W, bf16 e for the weight-gradient GEMM; fp32 accumulation) -- a restatement of
    what the kernels compute, for a tight check of every layer where bf16 rounding noise exceeds the physical signal."""
    L = len(om)
    dt_ = torch.float32 if faithful else torch.float64
    rnd = _bf16r if faithful else (lambda t: t)
    errs, hs = [], []
    for l in range(L):
        u = x if l == 0 else A[l - 1]
        h = rnd(O.act_fn(om[l]["act"], u.to(dt_)))
        pred = h @ rnd(om[l]["W"].to(dt_)).T
        errs.append((y if l == L - 1 else A[l]).to(dt_) - pred)
        hs.append(h)
    E = torch.stack([0.5 * (e.double() ** 2).sum() / B for e in errs])
    E = torch.stack([E[L - 1]] + [E[k] for k in range(1, L - 1)] + [E[0]])   # reference order [output, hidden.., first]
    gW = [(-(1.0 / B) * rnd(errs[l]).T @ hs[l]).double() for l in range(L)]
    return E, gW


def _c4_rows(B, n=64):
    """row subset that touches the first, a middle and the last 256-row block (different CTA pairs / tile rows)"""
    k = n // 4
    return torch.cat([torch.arange(0, k), torch.arange(B // 2 - k, B // 2 + k), torch.arange(B - k, B)])


def _c4_case(act, dt, name, report, faithful_gate):
    import jpc_b200 as J
    B, D, L, n = 4096, 2048, 8, 64
    om, sk, x, y = _case(D, D, L, D, act, B)
    J.set_compute_dtype("bf16")
    rows = _c4_rows(B, n)
    xs, ys = x[rows], y[rows]
    a0 = O.init_activities_with_ffwd(om, xs)
    with _with_env(JPCB_STAGED=7):
        acts, arena = _fused_step_raw(om, sk, x, y, "euler", 20 * dt, dt)
    with _with_env(JPCB_STAGED=0):
        acts0, arena0 = _fused_step_raw(om, sk, x, y, "euler", 20 * dt, dt)
    # 3. persistent schedule == phase-by-phase round-1 kernels
    for l in range(L - 1):
        assert rel(acts[l], acts0[l].double().cpu()) < 2e-5, f"{name}: activities[{l}] differ between the two schedules"
    gate_all(arena.gW, [g.double().cpu() for g in arena0.gW], 2e-3, f"c4.{name}.gW_vs_phase_path")
    gate_all([e for e in arena.E], [e.double().cpu() for e in arena0.E], 2e-3, f"c4.{name}.E_vs_phase_path")
    # 1. row-subset oracle
    dt_s = dt * n / B
    ref, steps = O.solve_inference(om, None, a0, ys, x=xs, solver="euler", max_t1=20 * dt_s, dt=dt_s, use_closed_form=True,
                                   event_atol=0.0, event_rtol=0.0)
    assert steps == 20
    gate_all([acts[l][rows.cuda()] for l in range(L - 1)], ref[:-1], TOLBF, f"c4.{name}.activities_rows")
    # 2. every layer, from the returned activities
    A = [a.float().cpu() for a in acts]
    Ef, gWf = _c4_recompute(om, A, x.float(), y.float(), B, faithful=True)
    E64, gW64 = _c4_recompute(om, [a.double() for a in A], x, y, B, faithful=False)
    report.append((name, "gW vs bf16-operand restatement", [round(rel(arena.gW[l], gWf[l]), 5) for l in range(L)]))
    report.append((name, "E  vs bf16-operand restatement", [round(rel(arena.E[l], Ef[l]), 6) for l in range(L)]))
    report.append((name, "gW vs fp64 recomputation", [round(rel(arena.gW[l], gW64[l]), 5) for l in range(L)]))
    report.append((name, "E  vs fp64 recomputation", [round(rel(arena.E[l], E64[l]), 6) for l in range(L)]))
    # what rounding the GEMM operands to bf16 does to each tensor, measured on the host for this very data: the noise floor
    # of ANY bf16 implementation of these quantities (the device's realisation differs -- tanh.approx, summation order --
    # but is the same process); the device may deviate from fp64 by 2e-2 of the tensor's own scale plus 3x that floor
    nE = [3.0 * abs(float(Ef[l]) - float(E64[l])) for l in range(L)]
    nW = [3.0 * float((gWf[l] - gW64[l]).abs().max()) for l in range(L)]
    report.append((name, "bf16 operand-rounding floor of gW (max abs) / max|gW|", [round(nW[l] / 3 / float(gW64[l].abs().max()), 5) for l in range(L)]))
    gate_all([e for e in arena.E], [e for e in E64], TOLBF, f"c4.{name}.E_fp64", noise=nE)
    gate_all(arena.gW, gW64, TOLBF, f"c4.{name}.gW_fp64", noise=nW)
    if faithful_gate:   # relu: the host restatement of the operand rounding is exact, so every layer must match it tightly
        gate_all([e for e in arena.E], [e for e in Ef], TOLBF, f"c4.{name}.E_faithful")
        gate_all(arena.gW, gWf, TOLBF, f"c4.{name}.gW_faithful")
    # the output layer is above the rounding noise at every step size
    assert rel(arena.E[0], E64[0]) < TOLBF and rel(arena.gW[L - 1], gW64[L - 1]) < TOLBF


def test_config4_wide_full_size_bf16():
    """2048-wide, 8 layers, batch 4096, bf16 operands, 20 Euler steps -- the whole fused train step on the GPU (the
    whole-solve launch of csrc/tc_staged.cuh), at the bench's step size AND the stress step dt = 0.1 B:
      1. samples are independent during inference: 64 rows spread over three row blocks of the full-batch
         result against the fp64 oracle run on those rows alone with the step rescaled by 64/4096 (the energy's 1/B),
         every activity tensor on its own scale;
      2. EVERY layer's energy and weight gradient against a host recomputation from the returned activities:
         * stress step (every layer moves by O(1)): fp64 with exact operands -- the reference's arithmetic -- at 2e-2 per
           tensor, for tanh (the bench's activation) and relu;
         * bench step: dt/B = 1.2e-4 moves the hidden activities by less than the bf16 rounding of the feed-forward pass
           they started from, so the hidden errors ARE operand-rounding noise (of ANY bf16 implementation: which operand
           entries cross a bf16 rounding boundary).  They are pinned by restating the rounding on the host (bf16 act(z),
           bf16 W, bf16 e, fp32 accumulate) -- exact for relu, where act and the rounding are exact operations; for tanh the
           device's tanh.approx decides differently from the host's tanh which entries cross, so there the hidden layers are
           pinned by (3) and the output layer (above the noise) by fp64;
      3. the one-launch solve against the round-1 kernels run phase by phase (JPCB_STAGED=0: an independent epilogue
         implementation, same operand roundings): activities to 2e-5, every gW / energy to 2e-3 -- a missed dependency or a
         stale read in the persistent schedule would show here.
    The per-layer numbers are printed (pytest -s) and kept under profiles/."""
    report = []
    try:
        _c4_case("tanh", 0.5, "tanh.bench", report, faithful_gate=False)
        _c4_case("tanh", 0.1 * 4096, "tanh.stress", report, faithful_gate=False)
        _c4_case("relu", 0.5, "relu.bench", report, faithful_gate=True)
        _c4_case("relu", 0.1 * 4096, "relu.stress", report, faithful_gate=True)
    finally:
        for r in report:
            print("c4 full size:", *r)


# ---- fp32 parity on the tensor cores: 3-way bf16 split operands, six products per GEMM ("f32x3") ------------
def test_bpc_f32x3_split_tensor_core_path():
    """The same bPC checks as the fp32 CUDA-core path, at the SAME 1e-5 tolerance, but every GEMM on tcgen05:
    ragged dims (padding of the split parts to the 64-wide k-block), biases, skips, tanh (accurate activations on this
    path), energy weights, fused inference with the stress step."""
    import jpc_b200 as J
    J.set_compute_dtype("f32x3")
    c = _bpc_case(10, 256, 4, 28, "tanh", 128, use_bias=True)
    _check_bpc(*c, TOL32)
    _check_bpc(*c, TOL32, af=0.5, ab=0.7)
    _check_bpc_infer(*c, TOL32, T=6, dt=0.1 * 128)
    c = _bpc_case(16, 128, 5, 72, "relu", 136, skips=True)
    _check_bpc(*c, TOL32)
    _check_bpc_infer(*c, TOL32, T=4, dt=0.5)
    c = _bpc_case(24, 200, 3, 40, "silu", 64)
    _check_bpc(*c, TOL32)


def test_config5_bpc_full_size_f32x3():
    """BASELINE config 5 at full size and fp32 parity on the tensor cores."""
    import jpc_b200 as J
    J.set_compute_dtype("f32x3")
    c = _bpc_case(10, 1024, 6, 784, "relu", 1024)
    _check_bpc(*c, TOL32)
    _check_bpc_infer(*c, TOL32, T=30, dt=0.5, kink_budget=3)


def test_pc_f32x3_split_tensor_core_path():
    """The PC entry points on the split path at the fp32 tolerance: every-entry-point check, fused Euler / Heun solves and
    train steps, cross-entropy, unsupervised, ePC, the adaptive solve, hybrid PC's regression."""
    import jpc_b200 as J
    J.set_compute_dtype("f32x3")
    om, sk, x, y = _case(24, 72, 4, 16, "tanh", 16, use_bias=True)
    _check_all(om, sk, x, y, TOL32)
    _check_all(om, sk, x, y, TOL32, gamma=2.0, lam=0.7, ad=0.05)
    _solve_case(om, sk, x, y, "euler", 4.0, 0.5, TOL32, perturbed=True)
    _solve_case(om, sk, x, y, "heun", 2.0, 0.8 * 16 / 10, TOL32, perturbed=True, ad=0.05)
    _check_fused(om, sk, x, y, "euler", 5.0, 0.5, TOL32)
    _check_fused(om, sk, x, y, "heun", 3.0, 0.5, TOL32)
    om, sk, x, y = _case(10, 130, 6, 10, "relu", 24, param_type="mupc", skips=True)   # padded dims, skips, relu variants
    _check_all(om, sk, x, y, TOL32, param_type="mupc")
    _check_fused(om, sk, x, y, "euler", 4.0, 0.3 * 24, TOL32, param_type="mupc")
    om, sk, x, y = _case(12, 40, 3, 8, "relu", 8, use_bias=True)
    _check_ce(om, sk, x, y, TOL32)
    _adaptive_case(om, sk, x, y, 2e-5, 4.0, None)
    _check_epc(om, sk, x, y, TOL32)
    om, y2, acts = _unsup_case(16, 40, 3, 8, "tanh", 8, True)
    _check_unsup(om, y2, acts, TOL32)
    # hybrid PC's layer-wise regression
    gen, amo = _hpc_models(16, 40, 3, 24, "relu", True)
    xh, yh = data(8, 16, 24, onehot_out=False)
    fgen, famo = f32_model(gen), f32_model(amo)
    aa = O.init_activities_with_amort(famo, fgen, yh)
    eq = [a + 0.3 * torch.randn(a.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64) for a in aa[1:]]
    for_amort = eq[::-1][1:]
    ag = O.compute_hpc_param_grads(famo, for_amort, aa, yh, xh)
    jg = J.compute_hpc_param_grads(to_gpu_model(amo), devs(for_amort), devs(aa), dev(yh), dev(xh))
    for l in range(3):
        assert rel(jg[l][1].weight, ag[l]["W"]) < TOL32 and rel(jg[l][1].bias, ag[l]["b"]) < TOL32


def test_config_sized_pc_f32x3():
    """A large fp32 PC network (the shape class of BASELINE config 4, at 1/4 width) at fp32 parity on the tensor cores."""
    import jpc_b200 as J
    J.set_compute_dtype("f32x3")
    om, sk, x, y = _case(512, 512, 6, 512, "relu", 512)
    _check_fused(om, sk, x, y, "euler", 10.0, 0.1 * 512, TOL32, kink_budget=4)


# ---- PCTrainer: the fixed-step train step as one replayed CUDA graph ---------------------------------------------
@pytest.mark.parametrize("dtype,opt_name", [("f32", "sgd"), ("f32", "adam"), ("bf16", "sgd"), ("f32x3", "sgd")])
def test_pc_trainer_graph_matches_make_pc_step(dtype, opt_name):
    """Three consecutive train steps on three different batches: the graph-replayed trainer and the functional
    `make_pc_step` must produce the same losses, energies, activities and parameters (same kernels, same order)."""
    import jpc_b200 as J
    J.set_compute_dtype(dtype)
    om, sk, x, y = _case(24, 64, 4, 16, "tanh", 32, use_bias=True, skips=False)
    gm = to_gpu_model(om)
    mk = (lambda: J.optim.sgd(0.05)) if opt_name == "sgd" else (lambda: J.optim.adam(1e-2))
    kw = dict(ode_solver=J.Heun(), max_t1=3, dt=0.5, stepsize_controller=J.ConstantStepSize())
    tr = J.PCTrainer(gm, mk(), batch_size=32, **kw)
    opt = mk()
    model, st = gm, opt.init((gm, None))
    g = torch.Generator().manual_seed(4)
    for it in range(3):
        xb = torch.randn(32, 24, generator=g).cuda()
        yb = torch.randn(32, 16, generator=g).cuda()
        ref = J.make_pc_step(model, opt, st, yb, input=xb, **kw)
        out = tr.step(yb, xb)
        torch.cuda.synchronize()
        # (loss / energy sums are accumulated with atomics: equal up to summation order)
        assert abs(float(out["loss"]) - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"]))
        assert rel(out["energies"], ref["energies"].double().cpu()) < 1e-6
        for p, q in zip(out["activities"], ref["activities"]):
            assert rel(p, q.double().cpu()) < 1e-6
        model, st = ref["model"], ref["opt_state"]
        for a, b in zip(J.nn.tree_leaves(tr.model), J.nn.tree_leaves(model)):
            assert rel(a, b.double().cpu()) < 1e-6
    assert rel(gm[0][1].weight, om[0]["W"]) < 1e-7   # the caller's model was not touched


# ---- whole-solve launch of the staged tcgen05 kernel (csrc/tc_staged.cuh): all T Euler steps in one dependency-tracked stream ----
def _with_env(**kw):
    import contextlib
    import os

    @contextlib.contextmanager
    def cm():
        old = {k: os.environ.get(k) for k in kw}
        os.environ.update({k: str(v) for k, v in kw.items()})
        try:
            yield
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return cm()


@pytest.mark.parametrize("shape", [
    # d_in, width, depth, d_out, act, B, max_t1, dt
    (72, 200, 3, 40, "relu", 136, 2.0, 0.5),       # ragged tiles, one row block
    (64, 264, 5, 48, "tanh", 600, 6.0, 0.5 * 600 / 64),  # three row blocks (the last partial), stress step, 12 steps > depth
    (128, 256, 2, 64, "tanh", 256, 1.0, 0.3),      # two layers: one ERR task + the hoisted-prediction GRAD task; clipped last step
    (96, 520, 6, 24, "relu", 264, 1.5, 0.5),       # fewer steps (3) than layers: every step inside the wavefront warm-up
])
@pytest.mark.parametrize("use_bias", [False, True])
def test_whole_solve_launch_bf16(shape, use_bias):
    """The one-launch solve must reproduce the fp64 oracle (2e-2) AND the phase-by-phase launches of the same kernels
    (same arithmetic, so equal up to the order of the energy atomics), from the feed-forward init (wavefront schedule, fused
    train step) and from a perturbed state (dense schedule, solve_inference)."""
    import jpc_b200 as J
    d_in, width, depth, d_out, act, B, max_t1, dt = shape
    J.set_compute_dtype("bf16")
    om, sk, x, y = _case(d_in, width, depth, d_out, act, B, use_bias=use_bias, onehot_out=False)   # (biases: the staged ERR tile adds them)
    with _with_env(JPCB_SOLVE_MIN_ROUNDS=0, JPCB_STAGED=7):
        n0 = J._lib.lib.jpcb_launch_count()
        _check_fused(om, sk, x, y, "euler", max_t1, dt, TOLBF)
        acts1, arena1 = _fused_step_raw(om, sk, x, y, "euler", max_t1, dt)
        n1 = J._lib.lib.jpcb_launch_count()
        _solve_case(om, sk, x, y, "euler", max_t1, dt, TOLBF, perturbed=True)
    with _with_env(JPCB_SOLVE_MIN_ROUNDS=0, JPCB_STAGED=3):
        n2 = J._lib.lib.jpcb_launch_count()
        acts2, arena2 = _fused_step_raw(om, sk, x, y, "euler", max_t1, dt)
        n3 = J._lib.lib.jpcb_launch_count()
    T = int(math.ceil(max_t1 / dt - 1e-6))
    assert (n3 - n2) - (n1 - n0) // 2 >= T, "the whole-solve path did not replace the per-phase launches"
    for p, q in zip(acts1[:-1], acts2[:-1]):
        assert rel(p, q.double().cpu()) < 1e-5
    for p, q in zip(arena1.gW, arena2.gW):
        assert rel(p, q.double().cpu()) < 1e-4   # (bf16 operand copies of nearly equal activities may round apart)
    assert rel(arena1.E, arena2.E.double().cpu()) < 1e-5
    # the storer's other way out of shared memory (coalesced st.global instead of TMA stores): same numbers
    with _with_env(JPCB_SOLVE_MIN_ROUNDS=0, JPCB_STAGED=7, JPCB_TS_STG=0 if os.environ.get("JPCB_TS_STG", "0") != "0" else 1):
        acts3, arena3 = _fused_step_raw(om, sk, x, y, "euler", max_t1, dt)
        _solve_case(om, sk, x, y, "euler", max_t1, dt, TOLBF, perturbed=True)
    with _with_env(JPCB_SOLVE_MIN_ROUNDS=0, JPCB_STAGED=3, JPCB_TS_STG=0 if os.environ.get("JPCB_TS_STG", "0") != "0" else 1):
        acts4, _ = _fused_step_raw(om, sk, x, y, "euler", max_t1, dt)
    for p, q, r4 in zip(acts1[:-1], acts3[:-1], acts4[:-1]):
        assert rel(p, q.double().cpu()) < 1e-6 and rel(p, r4.double().cpu()) < 1e-5
    for p, q in zip(arena1.gW, arena3.gW):
        assert rel(p, q.double().cpu()) < 1e-5


@pytest.mark.parametrize("dtype,tol", [("f32", TOL32), ("bf16", TOLBF), ("f32x3", TOL32)])
def test_fused_train_step_heun_cross_entropy_skips_biases(dtype, tol):
    """A corner of the fused train step the configs do not visit: Heun steps, cross-entropy output, identity skips, biases,
    in all three arithmetic modes -- loss, energies, activities and every parameter gradient against the oracle."""
    import jpc_b200 as J
    J.set_compute_dtype(dtype)
    d_in, width, depth, d_out, B = (64, 64, 5, 16, 32) if dtype != "f32" else (12, 12, 5, 6, 7)
    om = f32_model(O.make_mlp(0, d_in, width, depth, d_out, "tanh", use_bias=True))
    sk = O.make_skip_model(depth)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, d_in, generator=g, dtype=torch.float64).float().double()
    y = torch.nn.functional.one_hot(torch.randint(0, d_out, (B,), generator=g), d_out).double()
    ref = O.pc_train_step(om, sk, x, y, solver="heun", max_t1=2.0, dt=0.5, loss_id="ce")
    gm, gsk = to_gpu_model(om), _gpu_skip(sk)
    opt = J.optim.sgd(1.0)
    out = J.make_pc_step(gm, opt, opt.init((gm, gsk)), dev(y), input=dev(x), loss_id="ce", skip_model=gsk, ode_solver=J.Heun(),
                         max_t1=2.0, dt=0.5, stepsize_controller=J.ConstantStepSize())
    assert abs(float(out["loss"]) - float(ref["loss"])) <= tol * abs(float(ref["loss"]))
    assert rel(out["energies"], ref["energies"]) < tol
    for p, q in zip(out["activities"][:-1], ref["activities"][:-1]):
        assert rel(p[0], q) < tol
    gmax = max(float(gr["W"].abs().max()) for gr in ref["grads"])
    # The hidden layers sit near equilibrium after the feed-forward init: their errors e = z - prediction are 1e-3 of z, so a
    # 1e-7 (fp32) or 4e-3 (bf16 operands) relative rounding of the prediction is 1e-4 / several times e itself, and their
    # gradients (1e-2 of the output layer's) cannot carry the tolerance on their own scale in ANY arithmetic of that
    # precision -- gated relative to the gradient list's scale (the group arm of tests/helpers.py:gate_all)
    floor = lambda own: gmax
    for l, gr in enumerate(ref["grads"]):   # sgd(1.0): the parameter moved by exactly its gradient
        ulp = 1.2e-7 * float(gm[l][1].weight.abs().max())
        got = (gm[l][1].weight.double() - out["model"][l][1].weight.double()).cpu()
        assert float((got - gr["W"]).abs().max()) <= tol * floor(float(gr["W"].abs().max())) + ulp, l
        gotb = (gm[l][1].bias.double() - out["model"][l][1].bias.double()).cpu()
        assert float((gotb - gr["b"]).abs().max()) <= tol * floor(float(gr["b"].abs().max())) + ulp, l
    J.set_compute_dtype("f32")


def test_user_skip_at_layer_0_disables_the_wavefront_shortcut():
    """A skip at layer 0 enters the feed-forward init (jpc/_core/_init.py:69-78 honours any non-None skip) but not the energy
    (jpc/_core/_energies.py:111-126 only layers 1..L-2): err[0] = x != 0 at the init, so the hidden errors are NOT all zero
    there and the fused train step must evaluate every layer from the first step on."""
    import jpc_b200 as J
    d, depth, B = 16, 5, 8
    om = f32_model(O.make_mlp(0, d, d, depth, 6, "tanh", use_bias=True))
    sk = [True, True, True, True, False]
    x, y = data(B, d, 6, onehot_out=False)
    x, y = x.float().double(), y.float().double()
    ref = O.pc_train_step(om, sk, x, y, solver="euler", max_t1=2.0, dt=0.5)
    assert float(ref["energies"][-1]) > 1e-3          # (the first layer's error is large from the start)
    gm = to_gpu_model(om)
    gsk = [J.nn.Lambda(J.nn.Identity()) if f else None for f in sk]
    opt = J.optim.sgd(1.0)
    for dtype, tol in (("f32", TOL32), ("bf16", TOLBF)):
        J.set_compute_dtype(dtype)
        out = J.make_pc_step(gm, opt, opt.init((gm, gsk)), dev(y), input=dev(x), skip_model=gsk, ode_solver=J.Euler(), max_t1=2.0,
                             dt=0.5, stepsize_controller=J.ConstantStepSize())
        assert abs(float(out["loss"]) - float(ref["loss"])) <= tol * abs(float(ref["loss"]))
        assert rel(out["energies"], ref["energies"]) < tol
        for p, q in zip(out["activities"][:-1], ref["activities"][:-1]):
            assert rel(p[0], q) < tol
    J.set_compute_dtype("f32")


# ---- feed-forward chain of a deep narrow network as one launch (csrc/chain_ffwd.cuh) ----------------------------------
@pytest.mark.parametrize("shape", [
    # d_in, width, depth, d_out, act, B, bias, param_type, skips
    (12, 16, 8, 3, "relu", 5, True, "mupc", True),        # skips + scalings + biases, one partial CTA
    (33, 65, 7, 10, "tanh", 37, True, "sp", False),       # ragged dims (no 16-byte rows), several CTAs
    (784, 128, 12, 10, "gelu", 128, False, "ntp", True),  # config-3-shaped, shallower
])
def test_chain_ffwd_equals_layer_by_layer(shape):
    """The one-launch chain must give the oracle's feed-forward init (fp32 gate) and what the layer-by-layer launches give,
    including the loss and -- through the fused train step -- everything downstream of the operand copies it leaves."""
    import jpc_b200 as J
    d_in, width, depth, d_out, act, B, bias, ptype, skips = shape
    om, sk, x, y = _case(d_in, width, depth, d_out, act, B, use_bias=bias, param_type=ptype, skips=skips)
    gm, gsk = to_gpu_model(om), _gpu_skip(sk)
    ref = O.init_activities_with_ffwd(om, x, skips=sk, param_type=ptype)
    outs = {}
    for mode in (2, 0):
        with _with_env(JPCB_CHAIN_FFWD=mode):
            n0 = J._lib.lib.jpcb_launch_count()
            outs[mode] = J.init_activities_with_ffwd(gm, dev(x), skip_model=gsk, param_type=ptype)
            outs[mode, "n"] = J._lib.lib.jpcb_launch_count() - n0
            outs[mode, "step"] = _fused_step_raw(om, sk, x, y, "euler", 2.0, 0.5, param_type=ptype)
    assert outs[0, "n"] - outs[2, "n"] == depth - 1, (outs[0, "n"], outs[2, "n"])   # L launches became one
    for p, q, r in zip(outs[2], outs[0], ref):
        assert rel(p, r) < TOL32 and rel(p, q.double().cpu()) < 1e-5
    (a2, ar2), (a0, ar0) = outs[2, "step"], outs[0, "step"]
    assert abs(float(ar2.loss) - float(ar0.loss)) <= 1e-5 * abs(float(ar0.loss))
    for p, q in zip(a2[:-1], a0[:-1]):
        assert rel(p, q.double().cpu()) < 1e-5
    gate_all(ar2.gW, [g.double().cpu() for g in ar0.gW], 1e-4, "chain_ffwd.gW")


# ---- the reference's own numeric known answers, run THROUGH the CUDA path (reference tests/test_analytical.py) ----------
@pytest.mark.parametrize("dtype,B", [("f32", 2), ("f32x3", 8)])
def test_reference_kat_linear_equilibrium_through_cuda(dtype, B):
    """reference tests/test_analytical.py:162-224 (200 x update_pc_activities with optax.sgd(0.5 B) on a muPC linear net
    16 -> 128 -> 1, gamma = 0.01, lambda = gamma^2 width depth: the energy must reach the analytical equilibrium energy,
    rtol = atol = 1e-2 there) and :293-368 (300 iterations: energy and first-layer activities must match the analytical
    activity solution, 1e-2) -- with jpc_b200.update_pc_activities / pc_energy_fn / init_activities_with_ffwd doing the
    work, against the oracle's restatement of jpc/_core/_analytical.py (itself pinned in tests/test_oracle.py).  The
    reference's gates are asserted as they stand, then far tighter ones (the iteration has converged to fp32 accuracy).
    (f32x3 needs a batch that is a multiple of 8; the known answer holds for any batch.)"""
    import jpc_b200 as J
    J.set_compute_dtype(dtype)
    input_dim, width, depth, gamma = 16, 128, 2, 0.01
    model = f32_model(O.make_mlp(42, input_dim, width, depth, 1, "linear", param_type="mupc"))
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, input_dim, generator=g, dtype=torch.float64).float().double()
    y = torch.randn(B, 1, generator=g, dtype=torch.float64).float().double()
    lam = gamma ** 2 * width * depth
    kw = dict(param_type="mupc", gamma=gamma, output_energy_scaling=lam)
    gm, gx, gy = to_gpu_model(model), dev(x), dev(y)
    acts = J.init_activities_with_ffwd(gm, gx, param_type="mupc", gamma=gamma)
    opt = J.optim.sgd(0.5 * B)
    st = opt.init(acts)
    e200 = None
    for it in range(300):
        r = J.update_pc_activities((gm, None), acts, opt, st, gy, input=gx, **kw)
        acts, st = r["activities"], r["opt_state"]
        if it == 199:
            e200 = float(r["energy"])
    e300 = float(r["energy"])
    theory = float(O.linear_equilib_energy(model, None, x, y, **kw))
    assert math.isclose(theory, e200, rel_tol=1e-2, abs_tol=1e-2)          # reference gate (:224)
    assert math.isclose(theory, e200, rel_tol=1e-4)
    theory_acts = O.linear_activity_solution(model, None, x, y, **kw)
    theory_energy = float(J.pc_energy_fn((gm, None), devs(theory_acts), gy, x=gx, **kw))
    assert math.isclose(theory_energy, e300, rel_tol=1e-2, abs_tol=1e-2)   # reference gate (:362)
    assert math.isclose(theory_energy, e300, rel_tol=1e-4) and math.isclose(theory_energy, theory, rel_tol=1e-4)
    dist = float(torch.linalg.norm(acts[0].double().cpu() - theory_acts[0], dim=1).mean())
    assert dist < 1e-2                                                      # reference gate (:363-367)
    assert dist < 1e-4 * float(torch.linalg.norm(theory_acts[0], dim=1).mean())
    print(f"KAT[{dtype}]: theory {theory:.8f}  200 it {e200:.8f}  300 it {e300:.8f}  |z1 - z1*| {dist:.2e}")


def test_reference_kat_equilib_energy_is_loss_over_s_through_cuda():
    """reference tests/test_analytical.py:227-290: scalar output => F* = loss / s, with the loss of the feed-forward pass
    computed by jpcb_pc_ffwd_init (make_pc_step's `loss`), rtol = atol = 1e-5 as there."""
    import jpc_b200 as J
    input_dim, width, depth, gamma = 16, 64, 2, 0.01
    model = f32_model(O.make_mlp(3, input_dim, width, depth, 1, "linear", param_type="mupc"))
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, input_dim, generator=g, dtype=torch.float64).float().double()
    y = torch.randn(3, 1, generator=g, dtype=torch.float64).float().double()
    lam = gamma ** 2 * width * depth
    scal = O.get_param_scalings(model, x, None, "mupc", gamma)
    s = float(O.linear_equilib_rescaling([l["W"] for l in model], scal, lam)[0, 0])
    out = J.init_activities_with_ffwd(to_gpu_model(model), dev(x), param_type="mupc", gamma=gamma)[-1]
    loss = float(O.mse_loss(out.double().cpu(), y))
    F = float(O.linear_equilib_energy(model, None, x, y, param_type="mupc", gamma=gamma, output_energy_scaling=lam))
    assert math.isclose(F, loss / s, rel_tol=1e-5, abs_tol=1e-5)


def test_zz_group_arm_budget():
    """Runs last in this file: which tensors passed a parity gate only relative to the largest tensor of their group
    (tests/helpers.py: gate_all), and a bound on how many may.  They are gradients / energies of layers whose signal is
    orders of magnitude below their neighbours' (deep muPC layers at the edge of the error wavefront, config 3): in fp32 the
    rounding of the shared intermediate z - prediction dominates them in ANY fp32 implementation."""
    by_tag = {}
    for tag, i, own, grp in GROUP_ARM:
        by_tag.setdefault(tag, []).append((i, own, grp))
    for tag, v in sorted(by_tag.items()):
        worst = max(v, key=lambda t: t[1])
        print(f"group arm: {tag}: {len(v)} tensors, worst own-scale error {worst[1]:.2e} (tensor {worst[0]}), "
              f"largest group-scale error {max(t[2] for t in v):.2e}")
    print(f"group arm total: {len(GROUP_ARM)} of {N_GATED[0]} gated tensors")
    for tag, i, own, nz in NOISE_ARM:
        print(f"noise arm: {tag}[{i}]: error {own:.2e} of own max = {nz:.2f} x the stated floor")
    assert len(GROUP_ARM) <= GROUP_ARM_BUDGET, f"{len(GROUP_ARM)} tensors needed the group-relative arm (budget {GROUP_ARM_BUDGET})"

