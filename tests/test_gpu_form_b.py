"""Layers of the form act(W u + b) -- `nn.Sequential([nn.Linear, nn.Lambda(act)])`, the hand-built networks of the reference's
example notebooks (examples/jpc_from_scratch.ipynb cell 12, examples/discriminative_pc.ipynb cell 8; SURVEY Appendix A.1
"Form B") -- through every supervised PC entry point, against the fp64 oracle (autograd of the restated energy, whose
closed-form twin is checked in tests/test_oracle.py)."""
import pytest
import torch

from oracle import pc_oracle as O
from tests.helpers import data, dev, devs, f32_model, perturb, rel, to_gpu_model

pytestmark = pytest.mark.gpu

TOL = {"f32": 1e-5, "f32x3": 1e-5, "bf16": 2e-2}


@pytest.fixture(autouse=True)
def _f32():
    import jpc_b200 as J
    J.set_compute_dtype("f32")
    yield
    J.set_compute_dtype("f32")


def _lin(layer):
    """The Linear of a layer of either form."""
    return layer if hasattr(layer, "weight") else next(m for m in layer.layers if hasattr(m, "weight"))


def _case(act, dims=(16, 32, 4, 8), B=16, use_bias=True, seed=0):
    om = f32_model(O.make_form_b_mlp(seed, dims[0], dims[1], dims[2], dims[3], act, use_bias=use_bias))
    om[-1]["bare"] = True   # last layer: a bare nn.Linear, as in the notebooks
    x, y = data(B, dims[0], dims[3], onehot_out=False)
    return om, x.float().double(), y.float().double()


def _close(p, q, tol, scale):
    return float((p.double().cpu() - q).abs().max()) <= tol * scale


def _bf16(t):
    """Rounded to the nearest bf16-representable value (kept in fp64)."""
    return t.float().bfloat16().double()


def _entry_points(om, x, y, tol, bf16_exact=False):
    import jpc_b200 as J
    gm, gx, gy = to_gpu_model(om), dev(x), dev(y)
    a0 = O.init_activities_with_ffwd(om, x)
    for p, q in zip(J.init_activities_with_ffwd(gm, gx), a0):
        assert rel(p, q) < tol
    acts = [a.float().double() for a in perturb(a0)]
    if bf16_exact:
        # bf16 operands + a kinked activation AFTER the GEMM: act'(pre) is evaluated on a bf16 GEMM result, so a
        # pre-activation within operand rounding of zero takes the other branch and moves a whole gradient row by one
        # term (not a rounding-size effect).  With operands that ARE bf16 numbers (weights, input, activities) the forward
        # GEMMs are exact up to fp32 accumulation, the kinks sit where the fp64 oracle has them, and what is left to
        # compare is the smooth rounding of the backward operands -- the 2e-2 of this path.
        acts = [_bf16(a) for a in acts]
    gacts = devs(acts)
    E = O.pc_energy_fn(om, None, acts, y, x=x, record_layers=True)
    assert rel(J.pc_energy_fn((gm, None), gacts, gy, x=gx, record_layers=True), E) < tol
    F, g = O.compute_pc_activity_grad(om, None, acts, y, x=x)
    gF, gg = J.compute_pc_activity_grad((gm, None), gacts, gy, x=gx)
    assert abs(float(gF) - float(F)) <= tol * abs(float(F))
    gmax = max(float(t.abs().max()) for t in g)
    for p, q in zip(gg, g):
        assert _close(p, q, tol, max(float(q.abs().max()), 1e-3 * gmax))
    pg = O.compute_pc_param_grads(om, None, acts, y, x=x)
    gpg, _ = J.compute_pc_param_grads((gm, None), gacts, gy, x=gx)
    for l, q in enumerate(pg):
        assert _close(_lin(gpg[l]).weight, q["W"], tol, float(q["W"].abs().max())), l
        if q["b"] is not None:
            assert _close(_lin(gpg[l]).bias, q["b"], tol, float(q["b"].abs().max())), l


@pytest.mark.parametrize("act", O.ACT_FNS)
def test_form_b_every_activation_fp32(act):
    _entry_points(*_case(act), TOL["f32"])
    _entry_points(*_case(act, dims=(7, 13, 3, 5), B=6, use_bias=False), TOL["f32"])   # ragged dims, no biases


@pytest.mark.parametrize("dtype", ["bf16", "f32x3"])
@pytest.mark.parametrize("act", ["tanh", "relu"])
def test_form_b_on_the_tensor_core_paths(dtype, act):
    import jpc_b200 as J
    J.set_compute_dtype(dtype)
    # (f32x3 evaluates tanh accurately; a relu kink within rounding of zero may flip one entry's derivative: scale = own max)
    om, x, y = _case(act, dims=(64, 128, 4, 32), B=64)
    if dtype == "bf16":
        for layer in om:
            layer["W"] = _bf16(layer["W"])
        x = _bf16(x)
    _entry_points(om, x, y, TOL[dtype], bf16_exact=(dtype == "bf16"))


def test_form_b_with_an_input_activation_too():
    """Sequential([Lambda(act), Linear, Lambda(act)]): both activations of one layer (the descriptor composes them)."""
    om, x, y = _case("tanh")
    om[1]["act"] = "relu"
    om[2]["act"] = "silu"
    _entry_points(om, x, y, TOL["f32"])


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("solver", ["euler", "heun"])
def test_form_b_fixed_step_solve_and_fused_train_step(dtype, solver):
    import jpc_b200 as J
    J.set_compute_dtype(dtype)
    tol = TOL[dtype]
    om, x, y = _case("tanh", dims=(64, 128, 4, 32), B=64) if dtype == "bf16" else _case("tanh")
    gm, gx, gy = to_gpu_model(om), dev(x), dev(y)
    a0 = O.init_activities_with_ffwd(om, x)
    sol, n = O.solve_inference(om, None, a0, y, x=x, solver=solver, max_t1=3.0, dt=0.5, event_atol=0.0, event_rtol=0.0)
    S = J.Euler() if solver == "euler" else J.Heun()
    gsol = J.solve_inference((gm, None), devs(a0), gy, input=gx, solver=S, max_t1=3.0, dt=0.5, stepsize_controller=J.ConstantStepSize())
    for p, q in zip(gsol, sol):
        assert rel(p[0], q) < tol
    ref = O.pc_train_step(om, None, x, y, solver=solver, max_t1=3.0, dt=0.5)
    opt = J.optim.sgd(1.0)
    out = J.make_pc_step(gm, opt, opt.init((gm, None)), gy, input=gx, ode_solver=S, max_t1=3.0, dt=0.5,
                         stepsize_controller=J.ConstantStepSize())
    assert abs(float(out["loss"]) - float(ref["loss"])) <= tol * abs(float(ref["loss"]))
    assert rel(out["energies"], ref["energies"]) < tol
    for p, q in zip(out["activities"][:-1], ref["activities"][:-1]):
        assert rel(p[0], q) < tol
    gmax = max(float(g["W"].abs().max()) for g in ref["grads"])
    for l, g in enumerate(ref["grads"]):   # sgd(1.0): new parameter = old - gradient (read back through fp32 parameters:
        ulp = 1.2e-7 * float(_lin(gm[l]).weight.abs().max())   # the difference carries one rounding of the parameter itself)
        got = (_lin(gm[l]).weight.double() - _lin(out["model"][l]).weight.double()).cpu()
        assert float((got - g["W"]).abs().max()) <= tol * max(float(g["W"].abs().max()), 1e-2 * gmax) + ulp, l
        if g["b"] is not None:
            gotb = (_lin(gm[l]).bias.double() - _lin(out["model"][l]).bias.double()).cpu()
            assert float((gotb - g["b"]).abs().max()) <= tol * max(float(g["b"].abs().max()), 1e-2 * gmax) + ulp, l


def test_form_b_notebook_network_default_solver_learns():
    """The network of examples/discriminative_pc.ipynb cell 8 (784-300-300-10, relu after the Linear, bare last layer) through
    `make_pc_step` with the reference's default solver (Heun + PIDController): the adaptive solve matches the oracle's and the
    loss falls over a few steps."""
    import jpc_b200 as J
    om = f32_model(O.make_form_b_mlp(0, 784, 300, 3, 10, "relu", use_bias=True))
    om[-1]["bare"] = True
    x, y = data(64, 784, 10)
    x, y = x.float().double(), y.float().double()
    gm, gx, gy = to_gpu_model(om), dev(x), dev(y)
    a0 = O.init_activities_with_ffwd(om, x)
    ref, stats = O.solve_inference_adaptive(om, None, a0, y, x=x, max_t1=5.0)
    got = J.solve_inference((gm, None), devs(a0), gy, input=gx, max_t1=5)
    assert J.solve_inference.last_stats["accepted"] == stats["accepted"] and J.solve_inference.last_stats["rejected"] == stats["rejected"]
    for p, q in zip(got, ref):
        assert rel(p[0], q) < 2e-5
    opt = J.optim.adam(1e-3)
    model, st, losses = gm, opt.init((gm, None)), []
    for _ in range(6):
        out = J.make_pc_step(model, opt, st, gy, input=gx, max_t1=5)
        model, st = out["model"], out["opt_state"]
        losses.append(float(out["loss"]))
    assert losses[-1] < losses[0]
    assert type(model[0]).__name__ == "Sequential" and hasattr(model[0].layers[0], "weight") and hasattr(model[-1], "weight")


def test_form_b_limits_are_reported():
    import jpc_b200 as J
    om, x, y = _case("tanh")
    gm, gx, gy = to_gpu_model(om), dev(x), dev(y)
    with pytest.raises(NotImplementedError):   # the reference's scalings read model[0][1].weight: "sp" only in this form
        J.init_activities_with_ffwd(gm, gx, param_type="mupc")
    acts = [torch.zeros(16, d, device="cuda") for d in (16, 32, 32, 32, 8)]
    with pytest.raises((NotImplementedError, ValueError)):   # unsupervised mode
        J.pc_energy_fn((gm, None), acts, gy, x=None)
    with pytest.raises(NotImplementedError):   # bPC
        J.bpc_energy_fn(gm, gm, acts[1:], gy, x=gx)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_weight_decay_and_spectral_penalty_on_bare_linear_layers(dtype):
    """The reference applies both regularisers to layers that expose `.weight` directly (jpc/_core/_energies.py:127-143):
    the bare last Linear of the notebook networks, or every layer of a network built from bare Linears.  Energy value,
    parameter gradients and the update inside `make_pc_step` against the oracle's autograd; make_mlp models are untouched."""
    import jpc_b200 as J
    J.set_compute_dtype(dtype)
    tol = TOL[dtype]
    wd, sp = 0.3, 0.05
    dims, B = ((64, 128, 3, 24), 32) if dtype == "bf16" else ((10, 12, 3, 5), 4)
    cases = []
    om, x, y = _case("tanh", dims=dims, B=B)
    cases.append((om, x, y))
    lin = f32_model(O.make_mlp(3, dims[0], dims[1], dims[2], dims[3], "linear", use_bias=True))   # a network of bare Linears
    for layer in lin:
        layer["bare"] = True
    cases.append((lin, x, y))
    for om, x, y in cases:
        gm, gx, gy = to_gpu_model(om), dev(x), dev(y)
        acts = [a.float().double() for a in perturb(O.init_activities_with_ffwd(om, x))]
        if dtype == "bf16":
            acts = [_bf16(a) for a in acts]
        gacts = devs(acts)
        kw = dict(weight_decay=wd, spectral_penalty=sp)
        F = O.pc_energy_fn(om, None, acts, y, x=x, **kw)
        assert abs(float(J.pc_energy_fn((gm, None), gacts, gy, x=gx, **kw)) - float(F)) <= tol * abs(float(F))
        gF, _ = J.compute_pc_activity_grad((gm, None), gacts, gy, x=gx, **kw)
        assert abs(float(gF) - float(F)) <= tol * abs(float(F))
        assert abs(float(F) - float(O.pc_energy_fn(om, None, acts, y, x=x))) > 1e-3 * abs(float(F))   # (the terms are not negligible here)
        pg = O.compute_pc_param_grads(om, None, acts, y, x=x, **kw)
        gpg, _ = J.compute_pc_param_grads((gm, None), gacts, gy, x=gx, **kw)
        for l, q in enumerate(pg):
            assert _close(_lin(gpg[l]).weight, q["W"], tol, float(q["W"].abs().max())), l
            assert _close(_lin(gpg[l]).bias, q["b"], tol, float(q["b"].abs().max())), l
        # inside make_pc_step: sgd(1.0) moves every parameter by exactly its (regularised) gradient at the solution
        sol, _ = O.solve_inference(om, None, O.init_activities_with_ffwd(om, x), y, x=x, solver="euler", max_t1=1.0, dt=0.5,
                                   event_atol=0.0, event_rtol=0.0)
        ref = O.compute_pc_param_grads(om, None, sol, y, x=x, **kw)
        opt = J.optim.sgd(1.0)
        out = J.make_pc_step(gm, opt, opt.init((gm, None)), gy, input=gx, ode_solver=J.Euler(), max_t1=1.0, dt=0.5,
                             stepsize_controller=J.ConstantStepSize(), grad_norms=True, **kw)
        for l, g in enumerate(ref):
            ulp = 1.2e-7 * float(_lin(gm[l]).weight.abs().max())
            got = (_lin(gm[l]).weight.double() - _lin(out["model"][l]).weight.double()).cpu()
            assert float((got - g["W"]).abs().max()) <= tol * float(g["W"].abs().max()) + ulp, l
        assert out["model_grad_norms"] is not None
    # make_mlp layers are Sequential objects without `.weight`: both regularisers are no-ops, as in the reference
    mm = f32_model(O.make_mlp(0, dims[0], dims[1], dims[2], dims[3], "tanh"))
    gm, gx, gy = to_gpu_model(mm), dev(x), dev(y)
    ga = J.init_activities_with_ffwd(gm, gx)
    assert float(J.pc_energy_fn((gm, None), ga, gy, x=gx, weight_decay=wd, spectral_penalty=sp)) == pytest.approx(
        float(J.pc_energy_fn((gm, None), ga, gy, x=gx)), rel=1e-6)


def test_form_b_deep_network_through_the_chain_feed_forward():
    """Eight Form B layers: the one-launch feed-forward chain (csrc/chain_ffwd.cuh) applies the post-activation too, and
    the fused train step after it matches the oracle."""
    import jpc_b200 as J
    from tests.test_gpu_parity import _with_env
    om, x, y = _case("tanh", dims=(12, 24, 8, 5), B=6)
    gm, gx, gy = to_gpu_model(om), dev(x), dev(y)
    a0 = O.init_activities_with_ffwd(om, x)
    for mode in (2, 0):
        with _with_env(JPCB_CHAIN_FFWD=mode):
            for p, q in zip(J.init_activities_with_ffwd(gm, gx), a0):
                assert rel(p, q) < TOL["f32"], mode
    ref = O.pc_train_step(om, None, x, y, solver="euler", max_t1=2.0, dt=0.5)
    opt = J.optim.sgd(1.0)
    out = J.make_pc_step(gm, opt, opt.init((gm, None)), gy, input=gx, ode_solver=J.Euler(), max_t1=2.0, dt=0.5,
                         stepsize_controller=J.ConstantStepSize())
    assert abs(float(out["loss"]) - float(ref["loss"])) <= TOL["f32"] * abs(float(ref["loss"]))
    assert rel(out["energies"], ref["energies"]) < TOL["f32"]
    for p, q in zip(out["activities"][:-1], ref["activities"][:-1]):
        assert rel(p[0], q) < TOL["f32"]


def test_form_b_fused_train_step_f32x3():
    import jpc_b200 as J
    J.set_compute_dtype("f32x3")
    om, x, y = _case("tanh", dims=(64, 128, 4, 32), B=64)
    gm, gx, gy = to_gpu_model(om), dev(x), dev(y)
    ref = O.pc_train_step(om, None, x, y, solver="heun", max_t1=2.0, dt=0.5)
    opt = J.optim.sgd(1.0)
    out = J.make_pc_step(gm, opt, opt.init((gm, None)), gy, input=gx, ode_solver=J.Heun(), max_t1=2.0, dt=0.5,
                         stepsize_controller=J.ConstantStepSize())
    tol = TOL["f32x3"]
    assert abs(float(out["loss"]) - float(ref["loss"])) <= tol * abs(float(ref["loss"]))
    assert rel(out["energies"], ref["energies"]) < tol
    for p, q in zip(out["activities"][:-1], ref["activities"][:-1]):
        assert rel(p[0], q) < tol
    for l, g in enumerate(ref["grads"]):
        ulp = 1.2e-7 * float(_lin(gm[l]).weight.abs().max())
        got = (_lin(gm[l]).weight.double() - _lin(out["model"][l]).weight.double()).cpu()
        assert float((got - g["W"]).abs().max()) <= tol * float(g["W"].abs().max()) + ulp, l
