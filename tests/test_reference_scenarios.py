"""The reference's own test scenarios for the hot path, replayed against jpc_b200 with the reference's call
syntax (keyword names, argument order, result keys) -- a user switching `import jpc` -> `import jpc_b200 as jpc`,
`optax` -> `jpc.optim`, `diffrax` -> `jpc.{Heun, Euler, PIDController, ConstantStepSize}` keeps their code.
Scenario list and fixture shapes follow /root/reference/tests/{conftest,test_train,test_infer,test_updates,
test_init,test_energies,test_grads,test_utils,test_errors,test_test_functions}.py (batch 4, 10-20-20-5, relu,
no bias); PDM / analytical scenarios are out of scope.  Like the reference's, these
are smoke-level assertions (keys, shapes, finiteness); numerical parity is tests/test_gpu_parity.py's job."""
import pytest
import torch

pytestmark = pytest.mark.gpu

B, D_IN, HID, D_OUT, DEPTH = 4, 10, 20, 5, 3
SIZES = [D_IN, HID, HID, D_OUT]


@pytest.fixture(autouse=True)
def _f32():
    import jpc_b200 as jpc
    jpc.set_compute_dtype("f32")
    yield
    jpc.set_compute_dtype("f32")


@pytest.fixture
def jpc():
    import jpc_b200
    return jpc_b200


def _randn(*shape, seed=42):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).cuda()


@pytest.fixture
def simple_model(jpc):
    return jpc.make_mlp(key=42, input_dim=D_IN, width=HID, depth=DEPTH, output_dim=D_OUT, act_fn="relu", use_bias=False,
                        param_type="sp")


@pytest.fixture
def x():
    return _randn(B, D_IN, seed=1)


@pytest.fixture
def y():
    return _randn(B, D_OUT, seed=2)


@pytest.fixture
def y_onehot():
    g = torch.Generator().manual_seed(3)
    return torch.nn.functional.one_hot(torch.randint(0, D_OUT, (B,), generator=g), D_OUT).float().cuda()


def _finite(t):
    return bool(torch.isfinite(torch.as_tensor(t)).all())


# ---- test_train.py ------------------------------------------------------------------------------
@pytest.mark.parametrize("loss_id", ["mse", "ce"])
def test_make_pc_step_supervised(jpc, simple_model, x, y, y_onehot, loss_id):
    optim = jpc.optim.sgd(learning_rate=0.01)
    opt_state = optim.init((simple_model, None))
    result = jpc.make_pc_step(model=simple_model, optim=optim, opt_state=opt_state, output=y if loss_id == "mse" else y_onehot,
                              input=x, loss_id=loss_id, param_type="sp", ode_solver=jpc.Heun(), max_t1=5,
                              stepsize_controller=jpc.PIDController(rtol=1e-3, atol=1e-3))
    for k in ("model", "skip_model", "opt_state", "loss"):
        assert k in result
    assert len(result["model"]) == len(simple_model) and _finite(result["loss"])


def test_make_pc_step_unsupervised(jpc, simple_model, y):
    optim = jpc.optim.sgd(learning_rate=0.01)
    result = jpc.make_pc_step(model=simple_model, optim=optim, opt_state=optim.init((simple_model, None)), output=y, input=None,
                              loss_id="mse", param_type="sp", ode_solver=jpc.Heun(), max_t1=5,
                              stepsize_controller=jpc.PIDController(rtol=1e-3, atol=1e-3), key=42, layer_sizes=SIZES,
                              batch_size=B, sigma=0.05)
    assert "model" in result and "opt_state" in result and result["loss"] is None


def test_make_pc_step_with_regularization_and_metrics(jpc, simple_model, x, y):
    optim = jpc.optim.sgd(learning_rate=0.01)
    common = dict(model=simple_model, optim=optim, opt_state=optim.init((simple_model, None)), output=y, input=x, loss_id="mse",
                  param_type="sp", ode_solver=jpc.Heun(), max_t1=5, stepsize_controller=jpc.PIDController(rtol=1e-3, atol=1e-3))
    result = jpc.make_pc_step(weight_decay=0.01, spectral_penalty=0.01, activity_decay=0.01, **common)
    assert "model" in result and _finite(result["energies"])
    result = jpc.make_pc_step(record_activities=True, record_energies=True, activity_norms=True, param_norms=True,
                              grad_norms=True, calculate_accuracy=True, **common)
    for k in ("model", "activities", "energies", "activity_norms", "model_param_norms", "acc"):
        assert k in result
    assert result["activities"][0].ndim == 3 and tuple(result["energies"].shape) == (len(simple_model), 500)
    assert _finite(result["activity_norms"]) and _finite(result["model_grad_norms"]) and 0 <= float(result["acc"]) <= 100


def test_make_pc_step_invalid_input(jpc):
    model = jpc.make_mlp(42, 10, 20, 5, 3, "relu", False, "sp")
    optim = jpc.optim.sgd(learning_rate=0.01)
    with pytest.raises(ValueError):   # no input and no key / layer_sizes / batch_size
        jpc.make_pc_step(model=model, optim=optim, opt_state=optim.init((model, None)), output=_randn(4, 5), input=None,
                         loss_id="mse", param_type="sp", ode_solver=jpc.Heun(), max_t1=5,
                         stepsize_controller=jpc.PIDController(rtol=1e-3, atol=1e-3))


def _amortiser(jpc):
    return jpc.make_mlp(key=7, input_dim=D_OUT, width=HID, depth=DEPTH, output_dim=D_IN, act_fn="relu", use_bias=False,
                        param_type="sp")


@pytest.mark.parametrize("mode", ["supervised", "unsupervised", "recording"])
def test_make_hpc_step(jpc, simple_model, x, y, mode):
    generator, amortiser = simple_model, _amortiser(jpc)
    gen_optim, amort_optim = jpc.optim.sgd(learning_rate=0.01), jpc.optim.sgd(learning_rate=0.01)
    result = jpc.make_hpc_step(generator=generator, amortiser=amortiser, optims=(gen_optim, amort_optim),
                               opt_states=(gen_optim.init((generator, None)), amort_optim.init(amortiser)), output=y,
                               input=None if mode == "unsupervised" else x, ode_solver=jpc.Heun(), max_t1=5,
                               stepsize_controller=jpc.PIDController(rtol=1e-3, atol=1e-3),
                               record_activities=mode == "recording", record_energies=mode == "recording")
    for k in ("generator", "amortiser", "opt_states", "losses", "activities", "t_max", "energies"):
        assert k in result
    assert len(result["opt_states"]) == 2 and _finite(result["losses"][1])
    if mode == "recording":
        assert result["t_max"] is not None and _finite(result["energies"][1])


# ---- test_infer.py ------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", ["plain", "ce", "regularised", "record_iters", "record_every", "unsupervised"])
def test_solve_inference(jpc, simple_model, x, y, y_onehot, variant):
    if variant == "unsupervised":
        activities = jpc.init_activities_from_normal(key=42, layer_sizes=SIZES, mode="unsupervised", batch_size=B, sigma=0.05)
    else:
        activities = jpc.init_activities_with_ffwd(simple_model, x, param_type="sp")
    extra = {"regularised": dict(weight_decay=0.01, spectral_penalty=0.01, activity_decay=0.01),
             "record_iters": dict(record_iters=True), "record_every": dict(record_every=2)}.get(variant, {})
    solution = jpc.solve_inference(params=(simple_model, None), activities=activities, output=y_onehot if variant == "ce" else y,
                                   input=None if variant == "unsupervised" else x, loss_id="ce" if variant == "ce" else "mse",
                                   param_type="sp", solver=jpc.Heun(), max_t1=10,
                                   stepsize_controller=jpc.PIDController(rtol=1e-3, atol=1e-3), **extra)
    assert len(solution) == len(activities)
    for sol, act in zip(solution, activities):
        assert sol.ndim == 3 and tuple(sol.shape[-2:]) == tuple(act.shape)
    if variant == "record_every":
        assert solution[0].shape[0] == 5 + 1     # ts = arange(0, 10, 2), then t1


# ---- test_updates.py ----------------------------------------------------------------------------
@pytest.mark.parametrize("supervised", [True, False])
def test_update_pc_activities(jpc, simple_model, x, y, supervised):
    activities = (jpc.init_activities_with_ffwd(simple_model, x, param_type="sp") if supervised else
                  jpc.init_activities_from_normal(key=42, layer_sizes=SIZES, mode="unsupervised", batch_size=B, sigma=0.05))
    optim = jpc.optim.sgd(learning_rate=0.01)
    result = jpc.update_pc_activities(params=(simple_model, None), activities=activities, optim=optim,
                                      opt_state=optim.init(activities), output=y, input=x if supervised else None,
                                      loss_id="mse", param_type="sp")
    for k in ("energy", "activities", "grads", "opt_state"):
        assert k in result
    assert len(result["activities"]) == len(activities) and _finite(result["energy"])


@pytest.mark.parametrize("regularised", [False, True])
def test_update_pc_params(jpc, simple_model, x, y, regularised):
    activities = jpc.init_activities_with_ffwd(simple_model, x, param_type="sp")
    optim = jpc.optim.sgd(learning_rate=0.01)
    extra = dict(weight_decay=0.01, spectral_penalty=0.01, activity_decay=0.01) if regularised else {}
    result = jpc.update_pc_params(params=(simple_model, None), activities=activities, optim=optim,
                                  opt_state=optim.init((simple_model, None)), output=y, input=x, loss_id="mse",
                                  param_type="sp", **extra)
    for k in ("model", "skip_model", "grads", "opt_state"):
        assert k in result
    assert len(result["model"]) == len(simple_model)


def _bpc_models(jpc):
    nn = jpc.nn
    relu = jpc.get_act_fn("relu")
    top_down = jpc.make_mlp(key=42, input_dim=D_IN, width=HID, depth=DEPTH, output_dim=D_OUT, act_fn="relu", use_bias=False,
                            param_type="sp")
    bottom_up = [nn.Sequential([nn.Lambda(relu), nn.Linear(HID, D_IN, use_bias=False, key=100)])]
    for i in range(1, DEPTH - 1):
        bottom_up.append(nn.Sequential([nn.Lambda(relu), nn.Linear(HID, HID, use_bias=False, key=100 + i)]))
    bottom_up.append(nn.Sequential([nn.Lambda(relu), nn.Linear(D_OUT, HID, use_bias=False, key=100 + DEPTH)]))
    # hidden activities z_0 .. z_{H-1} plus the unused output slot.  (The reference's own smoke test passes only the
    # H hidden arrays, which makes its `activities[-2]` silently pick z_{H-2}; SURVEY row a8.  jpc_b200 asks for the
    # documented H + 1 entries and raises ValueError otherwise.)
    acts = jpc.init_activities_from_normal(key=456, layer_sizes=[HID] * (DEPTH - 1) + [D_OUT], mode="unsupervised",
                                           batch_size=B, sigma=0.05)
    return top_down, bottom_up, acts


def test_update_bpc_activities_and_params(jpc, x, y):
    top_down_model, bottom_up_model, activities = _bpc_models(jpc)
    optim = jpc.optim.sgd(learning_rate=0.01)
    result = jpc.update_bpc_activities(top_down_model=top_down_model, bottom_up_model=bottom_up_model, activities=activities,
                                       optim=optim, opt_state=optim.init(activities), output=y, input=x, param_type="sp")
    for k in ("energy", "activities", "grads", "opt_state"):
        assert k in result
    assert len(result["activities"]) == len(activities) and _finite(result["energy"])
    td_optim, bu_optim = jpc.optim.sgd(learning_rate=0.01), jpc.optim.sgd(learning_rate=0.01)
    result = jpc.update_bpc_params(top_down_model=top_down_model, bottom_up_model=bottom_up_model, activities=activities,
                                   top_down_optim=td_optim, bottom_up_optim=bu_optim,
                                   top_down_opt_state=td_optim.init(top_down_model),
                                   bottom_up_opt_state=bu_optim.init(bottom_up_model), output=y, input=x, param_type="sp")
    for k in ("models", "grads", "opt_states"):
        assert k in result
    assert len(result["models"]) == 2 and len(result["models"][0]) == len(top_down_model)
    assert len(result["models"][1]) == len(bottom_up_model)
    with pytest.raises(ValueError):
        jpc.bpc_energy_fn(top_down_model, bottom_up_model, activities[:-1], y, x=x)


# ---- test_init.py -------------------------------------------------------------------------------
def test_init_functions(jpc, simple_model, x, y):
    acts = jpc.init_activities_with_ffwd(simple_model, x, param_type="sp")
    assert [tuple(a.shape) for a in acts] == [(B, HID), (B, HID), (B, D_OUT)]
    mupc = jpc.make_mlp(key=42, input_dim=D_IN, width=HID, depth=DEPTH, output_dim=D_OUT, act_fn="relu", use_bias=False,
                        param_type="mupc")
    acts = jpc.init_activities_with_ffwd(mupc, x, param_type="mupc")
    assert len(acts) == len(mupc) and all(_finite(a) for a in acts)
    skip = jpc.make_skip_model(len(simple_model))
    acts = jpc.init_activities_with_ffwd(simple_model, x, skip_model=skip, param_type="sp")
    assert len(acts) == len(simple_model)
    sup = jpc.init_activities_from_normal(key=42, layer_sizes=SIZES, mode="supervised", batch_size=B, sigma=0.05)
    assert [tuple(a.shape) for a in sup] == [(B, HID), (B, HID), (B, D_OUT)]
    uns = jpc.init_activities_from_normal(key=42, layer_sizes=SIZES, mode="unsupervised", batch_size=B, sigma=0.05)
    assert [tuple(a.shape) for a in uns] == [(B, w) for w in SIZES]
    amort = jpc.init_activities_with_amort(_amortiser(jpc), simple_model, y)
    assert [tuple(a.shape) for a in amort] == [(B, D_IN), (B, HID), (B, HID), (B, D_OUT)]


# ---- test_energies.py / test_grads.py -----------------------------------------------------------
def test_energies_and_grads(jpc, simple_model, x, y, y_onehot):
    params = (simple_model, None)
    acts = jpc.init_activities_with_ffwd(simple_model, x, param_type="sp")
    e = jpc.pc_energy_fn(params=params, activities=acts, y=y, x=x, loss="mse", param_type="sp")
    assert _finite(e) and float(e) >= 0
    uns = jpc.init_activities_from_normal(key=42, layer_sizes=SIZES, mode="unsupervised", batch_size=B, sigma=0.05)
    assert _finite(jpc.pc_energy_fn(params=params, activities=uns, y=y, x=None, loss="mse", param_type="sp"))
    assert _finite(jpc.pc_energy_fn(params=params, activities=acts, y=y_onehot, x=x, loss="ce", param_type="sp"))
    assert _finite(jpc.pc_energy_fn(params=params, activities=acts, y=y, x=x, weight_decay=0.01, spectral_penalty=0.01,
                                    activity_decay=0.01))
    layers = jpc.pc_energy_fn(params=params, activities=acts, y=y, x=x, record_layers=True)
    assert tuple(layers.shape) == (len(simple_model),) and _finite(layers)
    for pt in ("sp", "mupc", "ntp"):
        m = jpc.make_mlp(key=42, input_dim=D_IN, width=HID, depth=DEPTH, output_dim=D_OUT, act_fn="relu", use_bias=False,
                         param_type=pt)
        a = jpc.init_activities_with_ffwd(m, x, param_type=pt)
        assert _finite(jpc.pc_energy_fn(params=(m, None), activities=a, y=y, x=x, param_type=pt))
    scal = jpc._get_param_scalings(model=simple_model, input=x, skip_model=None, param_type="sp")
    assert len(scal) == len(simple_model) and all(s == 1.0 for s in scal)
    # gradients
    t = torch.zeros(())
    args = (params, y, x, "mse", "sp", 0.0, 0.0, 0.0, None, None, None)
    neg = jpc.neg_pc_activity_grad(t, acts, args)
    energy, grads = jpc.compute_pc_activity_grad(params=params, activities=acts, y=y, x=x, loss_id="mse", param_type="sp")
    assert len(neg) == len(grads) == len(acts) and _finite(energy)
    for n, g in zip(neg, grads):
        assert torch.allclose(n, -g)
    energy, grads = jpc.compute_pc_activity_grad(params=params, activities=uns, y=y, x=None)
    assert len(grads) == len(uns)
    model_grads, skip_grads = jpc.compute_pc_param_grads(params=params, activities=acts, y=y, x=x, loss_id="mse", param_type="sp")
    assert len(model_grads) == len(simple_model) and skip_grads is None
    jpc.compute_pc_param_grads(params=params, activities=acts, y=y, x=x, weight_decay=0.01, spectral_penalty=0.01,
                               activity_decay=0.01)
    # hybrid PC energy / gradients with the reference's (shape-compatible) argument choice
    amortiser = _amortiser(jpc)
    amort = jpc.init_activities_with_amort(amortiser, simple_model, y)
    e = jpc.hpc_energy_fn(model=amortiser, equilib_activities=acts, amort_activities=amort, x=y, y=x)
    assert _finite(e) and float(e) >= 0
    g = jpc.compute_hpc_param_grads(model=amortiser, equilib_activities=acts, amort_activities=amort, x=y, y=x)
    assert len(g) == len(amortiser)
    # bPC
    td, bu, bacts = _bpc_models(jpc)
    assert _finite(jpc.bpc_energy_fn(top_down_model=td, bottom_up_model=bu, activities=bacts, y=y, x=x, param_type="sp"))
    energy, grads = jpc.compute_bpc_activity_grad(top_down_model=td, bottom_up_model=bu, activities=bacts, y=y, x=x)
    assert len(grads) == len(bacts) and _finite(energy)
    tdg, bug = jpc.compute_bpc_param_grads(top_down_model=td, bottom_up_model=bu, activities=bacts, y=y, x=x)
    assert len(tdg) == len(td) and len(bug) == len(bu)


# ---- ePC scenarios of test_init.py / test_energies.py / test_grads.py / test_updates.py ----------------------
def test_epc_scenarios(jpc, simple_model, x, y, y_onehot):
    errors = jpc.init_epc_errors(layer_sizes=SIZES, batch_size=x.shape[0], mode="supervised")
    assert [tuple(e.shape) for e in errors] == [(B, HID), (B, HID), (B, D_OUT)] and all(float(e.abs().sum()) == 0 for e in errors)
    assert len(jpc.init_epc_errors(layer_sizes=SIZES, batch_size=B, mode="unsupervised")) == len(SIZES)
    params = (simple_model, None)
    for loss, target in (("mse", y), ("ce", y_onehot)):
        energy = jpc.epc_energy_fn(params=params, errors=errors, y=target, x=x, loss=loss, param_type="sp")
        assert _finite(energy) and float(energy) >= 0
        energy, grads = jpc.compute_epc_error_grad(params=params, errors=errors, y=target, x=x, loss_id=loss, param_type="sp")
        assert len(grads) == len(errors) and _finite(energy)
        optim = jpc.optim.sgd(learning_rate=0.01)
        result = jpc.update_epc_errors(params=params, errors=errors, optim=optim, opt_state=optim.init(errors), output=target,
                                       input=x, loss_id=loss, param_type="sp")
        for k in ("energy", "errors", "grads", "opt_state"):
            assert k in result
        assert len(result["errors"]) == len(errors) and _finite(result["energy"])
    for pt in ("sp", "mupc", "ntp"):
        m = jpc.make_mlp(key=42, input_dim=D_IN, width=HID, depth=DEPTH, output_dim=D_OUT, act_fn="relu", use_bias=False,
                         param_type=pt)
        assert _finite(jpc.epc_energy_fn(params=(m, None), errors=errors, y=y, x=x, loss="mse", param_type=pt))
        model_grads, skip_grads = jpc.compute_epc_param_grads(params=(m, None), errors=errors, y=y, x=x, loss_id="mse", param_type=pt)
        assert len(model_grads) == len(m)
        optim = jpc.optim.sgd(learning_rate=0.01)
        result = jpc.update_epc_params(params=(m, None), errors=errors, optim=optim, opt_state=optim.init((m, None)), output=y,
                                       input=x, loss_id="mse", param_type=pt)
        for k in ("model", "skip_model", "grads", "opt_state"):
            assert k in result
    skip = jpc.make_skip_model(DEPTH)
    assert _finite(jpc.epc_energy_fn(params=(simple_model, skip), errors=errors, y=y, x=x, loss="mse", param_type="sp"))


# ---- test_utils.py / test_errors.py -------------------------------------------------------------
def test_utils_and_errors(jpc, simple_model, x, y):
    assert len(simple_model) == DEPTH
    for act in ("linear", "tanh", "hard_tanh", "relu", "leaky_relu", "gelu", "selu", "silu"):
        assert len(jpc.make_mlp(42, D_IN, HID, DEPTH, D_OUT, act, True, "sp")) == DEPTH
        assert callable(jpc.get_act_fn(act))
    with pytest.raises(ValueError):
        jpc.get_act_fn("not_an_activation")
    skip = jpc.make_skip_model(DEPTH)
    assert len(skip) == DEPTH and skip[0] is None and skip[-1] is None and skip[1] is not None
    preds, labels = _randn(B, D_OUT, seed=5), _randn(B, D_OUT, seed=6)
    assert _finite(jpc.mse_loss(preds, labels)) and float(jpc.mse_loss(preds, labels)) >= 0
    onehot = torch.eye(D_OUT).cuda()[:B]
    assert _finite(jpc.cross_entropy_loss(preds, onehot))
    assert 0 <= float(jpc.compute_accuracy(onehot, preds)) <= 100
    iters = [torch.zeros(100, B, HID).cuda()]
    iters[0][10, 0, 0] = 1.0
    assert int(jpc.get_t_max(iters)) == 9
    acts = jpc.init_activities_with_ffwd(simple_model, x, param_type="sp")
    norms = jpc.compute_activity_norms(acts)
    assert len(norms) == len(acts) and _finite(norms)
    model_norms, skip_norms = jpc.compute_param_norms((simple_model, None))
    assert len(model_norms) > 0 and skip_norms is None
    sol = jpc.solve_inference(params=(simple_model, None), activities=acts, output=y, input=x, solver=jpc.Heun(), max_t1=10,
                              stepsize_controller=jpc.PIDController(rtol=1e-3, atol=1e-3), record_iters=True)
    t_max = jpc.get_t_max(sol)
    energies = jpc.compute_infer_energies(params=(simple_model, None), activities_iters=sol, t_max=t_max, y=y, x=x, loss="mse",
                                          param_type="sp")
    assert energies.shape[0] == len(simple_model) and _finite(energies)
    for pt in ("sp", "mupc", "ntp"):
        jpc._check_param_type(pt)
    with pytest.raises(ValueError):
        jpc._check_param_type("invalid")


# ---- test_test_functions.py ---------------------------------------------------------------------
def test_evaluation_helpers(jpc, simple_model, x, y_onehot):
    loss, acc = jpc.test_discriminative_pc(model=simple_model, output=y_onehot, input=x, loss="ce", param_type="sp")
    assert _finite(loss) and 0 <= float(acc) <= 100
    loss, acc = jpc.test_discriminative_pc(model=simple_model, output=y_onehot, input=x, loss="ce", param_type="sp",
                                           skip_model=jpc.make_skip_model(len(simple_model)))
    assert _finite(loss) and _finite(acc)
    g = torch.Generator().manual_seed(9)
    input_onehot = torch.nn.functional.one_hot(torch.randint(0, D_IN, (B,), generator=g), D_IN).float().cuda()
    for loss_id in ("mse", "ce"):
        input_acc, output_preds = jpc.test_generative_pc(model=simple_model, output=y_onehot, input=input_onehot, key=1,
                                                         layer_sizes=SIZES, batch_size=B, loss_id=loss_id, param_type="sp",
                                                         sigma=0.05, ode_solver=jpc.Heun(), max_t1=10,
                                                         stepsize_controller=jpc.PIDController(rtol=1e-3, atol=1e-3))
        assert 0 <= float(input_acc) <= 100 and tuple(output_preds.shape) == (B, D_OUT)
    amort_acc, hpc_acc, gen_acc, output_preds = jpc.test_hpc(generator=simple_model, amortiser=_amortiser(jpc), output=y_onehot,
                                                             input=input_onehot, key=1, layer_sizes=SIZES, batch_size=B,
                                                             sigma=0.05, ode_solver=jpc.Heun(), max_t1=10,
                                                             stepsize_controller=jpc.PIDController(rtol=1e-3, atol=1e-3))
    assert all(0 <= float(a) <= 100 for a in (amort_acc, hpc_acc, gen_acc)) and tuple(output_preds.shape) == (B, D_OUT)


# ---- the reference README's two usage examples (README.md:62-145), batch dimension made explicit ------------------
def test_readme_quick_example_and_advanced_usage(jpc):
    x = torch.ones(1, 3).cuda()
    y = -x
    model = jpc.make_mlp(0, input_dim=3, width=50, depth=5, output_dim=3, act_fn="relu")
    optim = jpc.optim.adam(1e-3)
    opt_state = optim.init((model, None))
    result = jpc.make_pc_step(model=model, optim=optim, opt_state=opt_state, output=y, input=x)
    model, opt_state = result["model"], result["opt_state"]
    assert len(model) == 5 and _finite(result["loss"]) and opt_state["count"] == 1
    # advanced usage: explicit init -> inference loop -> parameter update (legacy spelling `update_params`)
    activity_optim = jpc.optim.sgd(0.5)
    activities = jpc.init_activities_with_ffwd(model=model, input=x)
    activity_opt_state = activity_optim.init(activities)
    for _ in range(len(model)):
        activity_update_result = jpc.update_pc_activities(params=(model, None), activities=activities, optim=activity_optim,
                                                          opt_state=activity_opt_state, output=y, input=x)
        activities = activity_update_result["activities"]
        activity_opt_state = activity_update_result["opt_state"]
    result = jpc.update_params(params=(model, None), activities=activities, optim=optim, opt_state=opt_state, output=y, input=x)
    assert len(result["model"]) == 5 and result["opt_state"]["count"] == 2
