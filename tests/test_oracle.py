"""Oracle self-checks (CPU): autograd restatement vs independent closed forms, structural facts of
SURVEY 8c(iii), and the reference's analytical known-answer tests (tests/test_analytical.py)."""
import math

import pytest
import torch

from oracle import pc_oracle as O

torch.manual_seed(0)


def _data(B, d_in, d_out, seed=1, dtype=torch.float64, onehot_out=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, d_in, generator=g, dtype=torch.float64).to(dtype)
    if onehot_out:
        idx = torch.randint(0, d_out, (B,), generator=g)
        y = torch.nn.functional.one_hot(idx, d_out).to(dtype)
    else:
        y = torch.randn(B, d_out, generator=g, dtype=torch.float64).to(dtype)
    return x, y


def _perturb(acts, seed=2, scale=0.3):
    g = torch.Generator().manual_seed(seed)
    return [a + scale * torch.randn(a.shape, generator=g, dtype=torch.float64).to(a.dtype) for a in acts]


@pytest.mark.parametrize("act", O.ACT_FNS)
@pytest.mark.parametrize("use_bias", [False, True])
def test_autograd_matches_closed_form(act, use_bias):
    model = O.make_mlp(0, 10, 20, 4, 5, act, use_bias=use_bias)
    x, y = _data(4, 10, 5)
    acts = _perturb(O.init_activities_with_ffwd(model, x))
    _, g_ad = O.compute_pc_activity_grad(model, None, acts, y, x=x)
    g_cf = O.closed_form_activity_grad(model, None, acts, y, x)
    for a, b in zip(g_ad, g_cf):
        assert O.max_rel_err(a, b) < 1e-12
    p_ad = O.compute_pc_param_grads(model, None, acts, y, x=x)
    p_cf = O.closed_form_param_grads(model, None, acts, y, x)
    for a, b in zip(p_ad, p_cf):
        assert O.max_rel_err(a["W"], b["W"]) < 1e-12
        if use_bias:
            assert O.max_rel_err(a["b"], b["b"]) < 1e-12


@pytest.mark.parametrize("act", O.ACT_FNS)
@pytest.mark.parametrize("use_bias", [False, True])
def test_form_b_layers_autograd_matches_closed_form(act, use_bias):
    """Layers of the form act(W u + b) (examples/jpc_from_scratch.ipynb cell 12): the autograd restatement and the closed
    forms of SURVEY Appendix A.2 "Form B" agree; the energy is the notebook's own (cell 10) up to the library's 1/2."""
    model = O.make_form_b_mlp(0, 10, 20, 4, 5, act, use_bias=use_bias)
    x, y = _data(4, 10, 5)
    acts0 = O.init_activities_with_ffwd(model, x)
    assert float(O.pc_energy_fn(model, None, acts0, y, x=x, record_layers=True)[1:].abs().max()) < 1e-25
    acts = _perturb(acts0)
    # the notebook's energy: sum of squared errors of z_l - act(W z_{l-1} + b), output layer linear, / batch (no 1/2)
    nb = 0.0
    for l, layer in enumerate(model):
        u = x if l == 0 else acts[l - 1]
        pre = u @ layer["W"].T + (0 if layer["b"] is None else layer["b"])
        pred = O.act_fn(act, pre) if l + 1 < len(model) else pre
        nb = nb + torch.sum(((y if l + 1 == len(model) else acts[l]) - pred) ** 2)
    assert abs(float(O.pc_energy_fn(model, None, acts, y, x=x)) - 0.5 * float(nb) / 4) < 1e-12
    _, g_ad = O.compute_pc_activity_grad(model, None, acts, y, x=x)
    g_cf = O.closed_form_activity_grad(model, None, acts, y, x)
    for a, b in zip(g_ad, g_cf):
        assert O.max_rel_err(a, b) < 1e-12
    p_ad = O.compute_pc_param_grads(model, None, acts, y, x=x)
    p_cf = O.closed_form_param_grads(model, None, acts, y, x)
    for a, b in zip(p_ad, p_cf):
        assert O.max_rel_err(a["W"], b["W"]) < 1e-12
        if use_bias:
            assert O.max_rel_err(a["b"], b["b"]) < 1e-12


@pytest.mark.parametrize("param_type", ["mupc", "ntp"])
def test_closed_form_skips_scalings_gamma(param_type):
    depth = 6
    model = O.make_mlp(0, 12, 16, depth, 3, "relu", param_type=param_type)
    skips = O.make_skip_model(depth)
    x, y = _data(5, 12, 3)
    kw = dict(param_type=param_type, gamma=0.5, output_energy_scaling=2.5)
    acts = _perturb(O.init_activities_with_ffwd(model, x, skips=skips, param_type=param_type, gamma=0.5))
    _, g_ad = O.compute_pc_activity_grad(model, skips, acts, y, x=x, activity_decay=0.1, **kw)
    g_cf = O.closed_form_activity_grad(model, skips, acts, y, x, activity_decay=0.1, **kw)
    for a, b in zip(g_ad, g_cf):
        assert O.max_rel_err(a, b) < 1e-12
    p_ad = O.compute_pc_param_grads(model, skips, acts, y, x=x, **kw)
    p_cf = O.closed_form_param_grads(model, skips, acts, y, x, **kw)
    for a, b in zip(p_ad, p_cf):
        assert O.max_rel_err(a["W"], b["W"]) < 1e-12


def test_structural_facts():
    model = O.make_mlp(0, 10, 20, 3, 5, "relu")  # reference conftest shapes (tests/conftest.py:16-87)
    x, y = _data(4, 10, 5)
    acts = O.init_activities_with_ffwd(model, x)
    E = O.pc_energy_fn(model, None, acts, y, x=x, record_layers=True)
    assert len(E) == len(model) and torch.all(E >= 0) and torch.all(torch.isfinite(E))
    # hidden errors are zero at the feed-forward init; only the output layer carries energy
    assert float(E[1:].abs().max()) < 1e-25
    tot = O.pc_energy_fn(model, None, acts, y, x=x)
    assert abs(float(tot) - float(E.sum())) < 1e-12
    # output energy at the ffwd init equals the mse loss
    assert abs(float(E[0]) - float(O.mse_loss(acts[-1], y))) < 1e-12
    _, g = O.compute_pc_activity_grad(model, None, acts, y, x=x)
    assert float(g[-1].abs().max()) == 0.0  # dummy slot
    assert float(g[0].abs().max()) == 0.0   # wavefront: only z_{L-1} moves on the first step (F7)
    assert float(g[1].abs().max()) > 0.0


def test_euler_equals_sgd_and_batch_sharding():
    model = O.make_mlp(0, 10, 20, 3, 5, "tanh", use_bias=True)
    x, y = _data(8, 10, 5)
    acts0 = O.init_activities_with_ffwd(model, x)
    sol, n = O.solve_inference(model, None, acts0, y, x=x, solver="euler", max_t1=5.0, dt=0.5)
    assert n == 10
    a = [t.clone() for t in acts0]
    for _ in range(10):
        _, g = O.compute_pc_activity_grad(model, None, a, y, x=x)
        a = O.sgd_update(a, g, 0.5)
    for p, q in zip(sol, a):
        assert O.max_rel_err(p, q) < 1e-13
    # closed-form vector field gives the same trajectory
    sol_cf, _ = O.solve_inference(model, None, acts0, y, x=x, solver="euler", max_t1=5.0, dt=0.5,
                                  use_closed_form=True)
    for p, q in zip(sol, sol_cf):
        assert O.max_rel_err(p, q) < 1e-12
    # heun differs from euler but both decrease the energy
    solh, nh = O.solve_inference(model, None, acts0, y, x=x, solver="heun", max_t1=5.0, dt=0.5)
    assert nh == 10
    e0 = float(O.pc_energy_fn(model, None, acts0, y, x=x))
    assert float(O.pc_energy_fn(model, None, sol, y, x=x)) < e0
    assert float(O.pc_energy_fn(model, None, solh, y, x=x)) < e0


def test_n_fixed_steps():
    assert O.n_fixed_steps(10, 0.5) == 20
    assert O.n_fixed_steps(25, 0.5) == 50
    assert O.n_fixed_steps(1, 0.3) == 4  # last step clipped


# ---- reference known-answer tests (reference tests/test_analytical.py) ------------------------
def _kat_setup():
    input_dim, width, depth, gamma, B = 16, 128, 2, 0.01, 2
    model = O.make_mlp(42, input_dim, width, depth, 1, "linear", param_type="mupc")
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, input_dim, generator=g, dtype=torch.float64)
    y = torch.randn(B, 1, generator=g, dtype=torch.float64)
    lam = gamma ** 2 * width * depth
    return model, x, y, gamma, lam, B


def test_kat_linear_equilib_energy_matches_inference():
    """reference tests/test_analytical.py:162-224 (rtol=atol=1e-2 there)."""
    model, x, y, gamma, lam, B = _kat_setup()
    kw = dict(param_type="mupc", gamma=gamma, output_energy_scaling=lam)
    theory = float(O.linear_equilib_energy(model, None, x, y, **kw))
    acts = O.init_activities_with_ffwd(model, x, param_type="mupc", gamma=gamma)
    energy = None
    for _ in range(200):
        energy, g = O.compute_pc_activity_grad(model, None, acts, y, x=x, **kw)
        acts = O.sgd_update(acts, g, 0.5 * B)
    assert math.isclose(theory, float(energy), rel_tol=1e-2, abs_tol=1e-2)
    assert math.isclose(theory, float(energy), rel_tol=1e-6)  # far tighter in fp64


def test_kat_linear_activity_solution_matches_inference():
    """reference tests/test_analytical.py:293-368."""
    model, x, y, gamma, lam, B = _kat_setup()
    kw = dict(param_type="mupc", gamma=gamma, output_energy_scaling=lam)
    acts = O.init_activities_with_ffwd(model, x, param_type="mupc", gamma=gamma)
    energy = None
    for _ in range(300):
        energy, g = O.compute_pc_activity_grad(model, None, acts, y, x=x, **kw)
        acts = O.sgd_update(acts, g, 0.5 * B)
    theory_acts = O.linear_activity_solution(model, None, x, y, **kw)
    theory_energy = float(O.pc_energy_fn(model, None, theory_acts, y, x=x, **kw))
    assert math.isclose(theory_energy, float(energy), rel_tol=1e-2, abs_tol=1e-2)
    assert float(torch.linalg.norm(acts[0] - theory_acts[0], dim=1).mean()) < 1e-2
    # and the two analytical answers agree with each other
    assert math.isclose(theory_energy, float(O.linear_equilib_energy(model, None, x, y, **kw)), rel_tol=1e-9)


def test_kat_equilib_energy_is_loss_over_s():
    """reference tests/test_analytical.py:227-290: scalar output => F* = loss / s."""
    input_dim, width, depth, gamma = 16, 64, 2, 0.01
    model = O.make_mlp(3, input_dim, width, depth, 1, "linear", param_type="mupc")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, input_dim, generator=g, dtype=torch.float64)
    y = torch.randn(3, 1, generator=g, dtype=torch.float64)
    lam = gamma ** 2 * width * depth
    scal = O.get_param_scalings(model, x, None, "mupc", gamma)
    s = float(O.linear_equilib_rescaling([l["W"] for l in model], scal, lam)[0, 0])
    out = O.init_activities_with_ffwd(model, x, param_type="mupc", gamma=gamma)[-1]
    loss = float(O.mse_loss(out, y))
    F = float(O.linear_equilib_energy(model, None, x, y, param_type="mupc", gamma=gamma,
                                      output_energy_scaling=lam))
    assert math.isclose(F, loss / s, rel_tol=1e-5, abs_tol=1e-5)


def test_kat_with_skips_deep_linear():
    """Skip branch of linear_equilib_energy against a long Euler run (deep linear mupc resnet)."""
    depth, width = 5, 24
    model = O.make_mlp(11, 8, width, depth, 2, "linear", param_type="mupc")
    skips = O.make_skip_model(depth)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 8, generator=g, dtype=torch.float64)
    y = torch.randn(3, 2, generator=g, dtype=torch.float64)
    kw = dict(param_type="mupc")
    theory = float(O.linear_equilib_energy(model, skips, x, y, **kw))
    sol = O.linear_activity_solution(model, skips, x, y, **kw)
    assert math.isclose(theory, float(O.pc_energy_fn(model, skips, sol, y, x=x, **kw)), rel_tol=1e-9)


# ---- bPC ------------------------------------------------------------------------------------
def test_bpc_energy_and_grads():
    H = 3
    td = O.make_mlp(0, 5, 12, H + 1, 9, "relu")
    bu = O.make_mlp(1, 9, 12, H + 1, 5, "tanh")[::-1]
    x, _ = _data(4, 5, 5)
    _, y = _data(4, 9, 9, seed=3, onehot_out=False)
    acts = _perturb(O.init_activities_with_ffwd(td, x))  # H+1 entries; last = dummy
    E = O.bpc_energy_fn(td, bu, acts, y, x=x, record_layers=True,
                        forward_energy_weight=0.7, backward_energy_weight=1.3)
    assert len(E) == 2 * (H + 1) and torch.all(E >= 0)
    tot, g = O.compute_bpc_activity_grad(td, bu, acts, y, x=x,
                                         forward_energy_weight=0.7, backward_energy_weight=1.3)
    assert abs(float(tot) - float(E.sum())) < 1e-12
    assert float(g[-1].abs().max()) == 0.0
    gtd, gbu = O.compute_bpc_param_grads(td, bu, acts, y, x=x)
    assert len(gtd) == H + 1 and len(gbu) == H + 1
    assert all(float(gi["W"].abs().max()) > 0 for gi in gtd + gbu)


def test_bpc_unclamped_input_matches_hand_derivative():
    """x = None (jpc/_core/_energies.py:324): z_0 is also the input of top_down[0] and the target of bottom_up[0].  The oracle's
    autograd gradient of z_0 against the four terms written out by hand (what the CUDA path assembles)."""
    H, d, af, ab = 3, 7, 0.7, 1.3
    td = O.make_mlp(0, d, d, H + 1, 4, "tanh", use_bias=True)
    bu = O.make_mlp(1, 4, d, H + 1, d, "tanh", use_bias=True)[::-1]
    td[0]["act"] = "tanh"   # (make_mlp leaves the first layer's input linear: give phi_0' something to do)
    x, _ = _data(5, d, d)
    _, y = _data(5, 4, 4, seed=3, onehot_out=False)
    acts = _perturb(O.init_activities_with_ffwd(td, x))
    kw = dict(forward_energy_weight=af, backward_energy_weight=ab)
    F, g = O.compute_bpc_activity_grad(td, bu, acts, y, x=None, **kw)
    Fc, gc = O.compute_bpc_activity_grad(td, bu, acts, y, x=acts[0], **kw)   # same energy, z_0's two extra roles frozen
    assert abs(float(F) - float(Fc)) < 1e-12
    B = y.shape[0]
    z0 = acts[0]
    e0 = z0 - (O.act_fn(td[0]["act"], z0) @ td[0]["W"].T + td[0]["b"])
    d0 = z0 - (O.act_fn(bu[0]["act"], z0) @ bu[0]["W"].T + bu[0]["b"])
    h = O.act_deriv(td[0]["act"], z0)                                        # phi_0'(z_0), elementwise
    extra = (-af * h * (e0 @ td[0]["W"]) + ab * d0) / B
    assert float((g[0] - (gc[0] + extra)).abs().max()) < 1e-12
    for a, b in zip(g[1:], gc[1:]):
        assert float((a - b).abs().max()) < 1e-14
    assert float(extra.abs().max()) > 1e-3


def test_errors():
    with pytest.raises(ValueError):
        O.check_param_type("bogus")
    with pytest.raises(ValueError):
        O.act_fn("bogus", torch.zeros(1))


def test_adaptive_heun_matches_fine_fixed_step():
    """The PID-controlled Heun solve agrees with a very fine fixed-step solve to the controller's tolerance,
    takes far fewer steps, and shrinks/accepts as the I-controller prescribes."""
    om = O.make_mlp(0, 10, 20, 3, 5, "tanh", use_bias=True)
    x, y = _data(6, 10, 5)
    a0 = O.init_activities_with_ffwd(om, x)
    fine, _ = O.solve_inference(om, None, a0, y, x=x, solver="heun", max_t1=4.0, dt=4.0 / 4000, event_atol=0.0, event_rtol=0.0)
    ada, st = O.solve_inference_adaptive(om, None, a0, y, x=x, max_t1=4.0, dt0=None, rtol=1e-3, atol=1e-3)
    assert st["accepted"] + st["rejected"] < 400 and st["accepted"] >= 2
    assert abs(sum(st["dts"]) - 4.0) < 1e-9
    for p, q in zip(ada[:-1], fine[:-1]):
        assert float((p - q).abs().max()) < 5e-2 * max(float(q.abs().max()), 1.0)
    # a huge first step must be rejected and shrunk by at least factormin
    _, st2 = O.solve_inference_adaptive(om, None, a0, y, x=x, max_t1=4.0, dt0=1e4, rtol=1e-6, atol=1e-6)
    assert st2["rejected"] >= 1


def test_saveat_bookkeeping_matches_oracle_interpolation():
    """jpc_b200._core._SaveAt (host-side row bookkeeping of record_iters / record_every) against the oracle's
    `saveat_ts` on a synthetic trajectory -- no device work involved."""
    import types
    from jpc_b200._core import _SaveAt
    g = torch.Generator().manual_seed(0)
    ts_steps = [0.0, 0.4, 0.8, 1.2, 1.5]
    traj = [(t, [torch.randn(3, 4, generator=g), torch.randn(3, 2, generator=g)]) for t in ts_steps]
    net = types.SimpleNamespace(B=3, L=2, true_adims=[4, 2], dev=torch.device("cpu"))
    save = _SaveAt(net, 1.5, False, 0.5)     # ts = [0, 0.5, 1.0]
    for (t0, y0), (t1, y1) in zip(traj[:-1], traj[1:]):
        save.step(t0, y0, t1, y1)
    ys = save.finish(traj[-1][1])
    want = O.saveat_ts(traj, [0.0, 0.5, 1.0])
    for l in range(2):
        assert tuple(ys[l].shape) == (4,) + tuple(traj[0][1][l].shape)
        for i in range(3):
            assert torch.allclose(ys[l][i], want[i][l], atol=1e-6)
        assert torch.equal(ys[l][3], traj[-1][1][l])
    save = _SaveAt(net, 1.5, True, None)     # every step, inf-padded to diffrax's max_steps
    for (t0, y0), (t1, y1) in zip(traj[:-1], traj[1:]):
        save.step(t0, y0, t1, y1)
    ys = save.finish(traj[-1][1])
    assert ys[0].shape[0] == 4096 and torch.equal(ys[0][3], traj[-1][1][0]) and bool(torch.isinf(ys[0][4:]).all())
    from jpc_b200._utils import get_t_max
    assert int(get_t_max(ys)) == 3
    # an early stop (steady-state event after the first step, t = 0.4): diffrax writes the t1 row at the current save
    # index -- right after the one `ts` row it filled -- and the rows beyond stay inf
    save = _SaveAt(net, 1.5, False, 0.5)
    save.step(*traj[0], *traj[1])
    ys = save.finish(traj[1][1])
    assert torch.equal(ys[0][1], traj[1][1][0]) and bool(torch.isinf(ys[0][2:]).all()) and bool(torch.isfinite(ys[0][0]).all())


def test_weight_regularisers_apply_to_bare_linear_layers_only():
    """jpc/_core/_energies.py:127-143: weight_decay / spectral_penalty see only layers that expose `.weight` (bare Linear);
    their gradient is wd W + 2 sp W (W^T W - I); make_mlp models are untouched."""
    model = O.make_form_b_mlp(0, 10, 12, 3, 5, "tanh", use_bias=True)
    model[-1]["bare"] = True
    x, y = _data(4, 10, 5)
    acts = _perturb(O.init_activities_with_ffwd(model, x))
    wd, sp = 0.3, 0.05
    F0 = O.pc_energy_fn(model, None, acts, y, x=x)
    F1 = O.pc_energy_fn(model, None, acts, y, x=x, weight_decay=wd, spectral_penalty=sp)
    W = model[-1]["W"]
    S = W.T @ W - torch.eye(W.shape[1], dtype=W.dtype)
    assert abs(float(F1 - F0) - float(0.5 * wd * (W ** 2).sum() + 0.5 * sp * (S ** 2).sum())) < 1e-12
    g0 = O.compute_pc_param_grads(model, None, acts, y, x=x)
    g1 = O.compute_pc_param_grads(model, None, acts, y, x=x, weight_decay=wd, spectral_penalty=sp)
    assert O.max_rel_err(g1[-1]["W"] - g0[-1]["W"], wd * W + 2 * sp * W @ S) < 1e-12
    for a, b in zip(g0[:-1], g1[:-1]):
        assert torch.equal(a["W"], b["W"])
    mlp = O.make_mlp(0, 10, 12, 3, 5, "tanh")
    a2 = _perturb(O.init_activities_with_ffwd(mlp, x))
    assert float(O.pc_energy_fn(mlp, None, a2, y, x=x, weight_decay=wd, spectral_penalty=sp)) == float(O.pc_energy_fn(mlp, None, a2, y, x=x))


def test_optimiser_restatements_match_torch_optim():
    """optax is absent from the image, so `sgd_update` / `adam_update` restate its published rules.
    torch.optim.SGD / Adam are independent implementations of the same published algorithms (Adam with bias correction, eps
    added outside the square root): six steps on the same gradients must agree to rounding in fp64.  This pins the rule the
    fused CUDA update is tested against (`test_fused_optimiser_matches_optax_protocol`) to something outside this repository."""
    g = torch.Generator().manual_seed(5)
    shapes = [(7, 5), (7,), (3, 7)]
    p0 = [torch.randn(s, generator=g, dtype=torch.float64) for s in shapes]
    grads = [[torch.randn(s, generator=g, dtype=torch.float64) * (10.0 ** (k % 3 - 1)) for s in shapes] for k in range(6)]
    for name in ("sgd", "adam"):
        tp = [p.clone().requires_grad_(True) for p in p0]
        opt = torch.optim.SGD(tp, lr=0.05) if name == "sgd" else torch.optim.Adam(tp, lr=0.01, betas=(0.9, 0.999), eps=1e-8)
        mine, state = [p.clone() for p in p0], None
        for gs in grads:
            for t, gr in zip(tp, gs):
                t.grad = gr.clone()
            opt.step()
            if name == "sgd":
                mine = O.sgd_update(mine, gs, 0.05)
            else:
                mine, state = O.adam_update(mine, gs, state, 0.01)
            for a, b in zip(mine, tp):
                assert float((a - b.detach()).abs().max()) <= 1e-13 * max(1.0, float(b.detach().abs().max()))


def test_integrators_converge_to_the_matrix_exponential_solution():
    """diffrax is absent, so the Euler / Heun tableaux and the adaptive controller are restatements.
    For a LINEAR network the inference ODE is dz/dt = -(H z - c), whose exact flow is z* + expm(-H t)(z_0 - z*)
    (scipy.linalg.expm: independent of this repository).  Against it: Euler's error halves and Heun's quarters when the step
    halves (orders 1 and 2 -- a wrong Heun tableau shows up as order 1), and the adaptive Heun + PID solve lands within its
    tolerance."""
    import numpy as np
    from scipy.linalg import expm
    model = O.make_mlp(0, 4, 5, 3, 3, "linear", use_bias=True)
    x, y = _data(2, 4, 3, onehot_out=False)
    acts = _perturb(O.init_activities_with_ffwd(model, x))
    shapes = [a.shape for a in acts]
    sizes = [a.numel() for a in acts]

    def unflat(v):
        out, o = [], 0
        for s, n in zip(shapes, sizes):
            out.append(v[o:o + n].reshape(s))
            o += n
        return out

    def grad(v):
        return torch.cat([g.reshape(-1) for g in O.compute_pc_activity_grad(model, None, unflat(v), y, x=x)[1]])

    n = sum(sizes)
    z0 = torch.cat([a.reshape(-1) for a in acts])
    g0 = grad(torch.zeros(n, dtype=torch.float64))                       # = -c
    H = torch.stack([grad(torch.eye(n, dtype=torch.float64)[i]) - g0 for i in range(n)], dim=1)
    assert float((grad(z0) - (H @ z0 + g0)).abs().max()) < 1e-12          # the vector field is affine
    live = [i for i in range(n) if float(H[i].abs().sum()) > 0]           # (the dummy output slot has no dynamics)
    Hl, cl = H[live][:, live].numpy(), -(g0[live] + H[live][:, [i for i in range(n) if i not in live]] @ z0[[i for i in range(n) if i not in live]]).numpy()
    t1 = 4.0
    zstar = np.linalg.solve(Hl, cl)
    exact = zstar + expm(-Hl * t1) @ (z0[live].numpy() - zstar)

    def err(solver, dt):
        sol, _ = O.solve_inference(model, None, acts, y, x=x, solver=solver, max_t1=t1, dt=dt, event_atol=0.0, event_rtol=0.0)
        v = torch.cat([a.reshape(-1) for a in sol])[live].numpy()
        return float(np.abs(v - exact).max())

    e1, e2 = err("euler", 0.1), err("euler", 0.05)
    h1, h2 = err("heun", 0.1), err("heun", 0.05)
    assert 1.8 < e1 / e2 < 2.2, (e1, e2)        # first order
    assert 3.6 < h1 / h2 < 4.4, (h1, h2)        # second order
    assert h1 < 0.1 * e1
    sol, stats = O.solve_inference_adaptive(model, None, acts, y, x=x, max_t1=t1)
    v = torch.cat([a.reshape(-1) for a in sol])[live].numpy()
    if not stats.get("event", False):          # ran to t1: within a few tolerances of the exact flow
        assert float(np.abs(v - exact).max()) < 2e-2 * max(1.0, float(np.abs(exact).max()))
