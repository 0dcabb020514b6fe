"""Hybrid predictive coding (Tschantz et al. 2023) on the B200 path: the reference's `make_hpc_step`
(jpc/_train.py:270-441), `init_activities_with_amort` (jpc/_core/_init.py:126-160), `hpc_energy_fn`
(jpc/_core/_energies.py:161-243), `compute_hpc_param_grads` (jpc/_core/_grads.py:502-541) and `test_hpc`
(jpc/_test.py:180-271).  A generator (an ordinary PC network: the inference and parameter-gradient entry points)
plus an amortiser whose energy is a layer-wise regression of the generator's equilibrated activities
(`jpcb_hpc_param_grads`).  This module only maps the reference's list conventions onto those calls."""
from __future__ import annotations

import torch

from . import _lib
from ._core import (_Net, _grads_pytree, compute_pc_param_grads, init_activities_from_normal, init_activities_with_ffwd,
                    solve_inference)
from ._utils import compute_accuracy, compute_infer_energies, get_t_max
from .optim import update_and_apply


def init_activities_with_amort(amortiser, generator, input):
    """jpc/_core/_init.py:126-160: the amortiser's feed-forward guesses in the generator's order (reversed), followed
    by the generator's prediction of its target from the guess of its last hidden layer (a dummy slot)."""
    acts = init_activities_with_ffwd(amortiser, input)[::-1]
    acts.append(init_activities_with_ffwd([generator[-1]], acts[-1])[0])
    return acts


def _regression(model, equilib_activities, amort_activities, x, y, want_grads, want_layers):
    """The index bookkeeping of hpc_energy_fn (jpc/_core/_energies.py:228-243): layer l of the amortiser reads its own
    guess of layer l-1 (x for l = 0) and is regressed onto equilib_activities[l] (y, if given, for the last layer)."""
    L = len(model)
    guesses = list(amort_activities)[::-1][1:]          # [a_0 .. a_{L-1}]: the amortiser's own layer outputs
    if len(guesses) < L - 1:
        raise ValueError(f"expected at least {L} amortised activities, got {len(amort_activities)}")
    U = [x] + [guesses[l - 1] for l in range(1, L)]
    T = [equilib_activities[l] for l in range(L - 1)] + [y if y is not None else equilib_activities[-1]]
    net = _Net(model, None, x, T[-1], param_type="sp", loss_id="mse", gamma=None, output_energy_scaling=None,
               activity_decay=0.0, weight_decay=0.0, spectral_penalty=0.0)
    for l in range(L):
        for name, t, w in (("input", U[l], net.true_dims[l]), ("target", T[l], net.true_dims[l + 1])):
            if tuple(t.shape) != (net.B, w) or t.device != net.dev or t.dtype != torch.float32:
                raise ValueError(f"amortiser layer {l}: {name} has shape {tuple(t.shape)}, expected {(net.B, w)} (fp32, same device)")
    P = torch.nn.functional.pad
    Up = [P(U[l], (0, net.dims[l] - U[l].shape[1])).contiguous() for l in range(L)]
    Tp = [P(T[l], (0, net.dims[l + 1] - T[l].shape[1])).contiguous() for l in range(L)]
    E = torch.empty(L, dtype=torch.float32, device=net.dev)
    gW = [torch.empty(s, dtype=torch.float32, device=net.dev) for s in net.w_shapes] if want_grads else None
    gb = [torch.empty(s, dtype=torch.float32, device=net.dev) for s in net.b_shapes] if want_grads and net.has_bias else None
    ws = net.workspace()
    _lib.check(_lib.lib.jpcb_hpc_param_grads(net.desc, net.W, net.b, _lib.ptr_array(Up), _lib.ptr_array(Tp),
                                             _lib.ptr_array(gW) if gW else None, _lib.ptr_array(gb) if gb else None,
                                             _lib.ptr(E), _lib.ptr(ws), ws.numel(), net.stream()))
    grads = _grads_pytree(net, *net.out_grads(gW, gb))[0] if want_grads else None
    return (E if want_layers else E.sum()), grads


def hpc_energy_fn(model, equilib_activities, amort_activities, x, y=None, record_layers=False):
    """jpc/_core/_energies.py:161-243 (total, or per layer [last, 1..L-2, first], each / batch size; no 1/2)."""
    return _regression(model, equilib_activities, amort_activities, x, y, False, record_layers)[0]


def compute_hpc_param_grads(model, equilib_activities, amort_activities, x, y=None):
    """jpc/_core/_grads.py:502-541: gradient of hpc_energy_fn wrt the amortiser, with the amortiser's structure."""
    return _regression(model, equilib_activities, amort_activities, x, y, True, False)[1]


def make_hpc_step(generator, amortiser, optims, opt_states, output, *, input=None, ode_solver=None, max_t1=300, dt=None,
                  stepsize_controller=None, record_activities=False, record_energies=False):
    """jpc/_train.py:270-441.  Same arguments, same 7-key result dict."""
    if record_energies:
        record_activities = True
    gen_params = (generator, None)
    gen_optim, amort_optim = optims
    gen_opt_state, amort_opt_state = opt_states
    if input is not None:
        gen_activities = init_activities_with_ffwd(generator, input, skip_model=None)
        gen_loss = torch.mean((output - gen_activities[-1]) ** 2)
    else:
        gen_loss = None
    amort_activities = init_activities_with_amort(amortiser, generator, output)
    equilib = solve_inference(gen_params, amort_activities[1:] if input is not None else amort_activities, output,
                              input=input, solver=ode_solver, max_t1=max_t1, dt=dt, stepsize_controller=stepsize_controller,
                              record_iters=record_activities)
    t_max = get_t_max(equilib) if record_activities else None
    gen_energies = compute_infer_energies(gen_params, equilib, t_max, output, x=input) if record_energies else None
    row = t_max if record_activities else 0
    final = [a[row].contiguous() for a in equilib]
    for_amort = final[::-1][1:]                        # without the generator's dummy target prediction
    amort_loss = (torch.mean((input - amort_activities[0]) ** 2) if input is not None
                  else torch.mean((for_amort[-1] - amort_activities[0]) ** 2))
    amort_energies = hpc_energy_fn(amortiser, for_amort, amort_activities, output, input) if record_energies else None
    gen_grads = compute_pc_param_grads(gen_params, final, output, x=input)
    amort_grads = compute_hpc_param_grads(amortiser, for_amort, amort_activities, output, input)
    (generator, _), gen_opt_state = update_and_apply(gen_optim, gen_grads, gen_opt_state, gen_params)
    amortiser, amort_opt_state = update_and_apply(amort_optim, amort_grads, amort_opt_state, amortiser)
    return {
        "generator": generator,
        "amortiser": amortiser,
        "opt_states": (gen_opt_state, amort_opt_state),
        "activities": (amort_activities, equilib),
        "t_max": t_max,
        "losses": (gen_loss, amort_loss),
        "energies": (gen_energies, amort_energies),
    }


def test_hpc(generator, amortiser, output, input, key, layer_sizes, batch_size, sigma=0.05, ode_solver=None, max_t1=500,
             dt=None, stepsize_controller=None):
    """jpc/_test.py:180-271: input accuracies of the amortiser, of the hybrid (amortised init + inference) and of the
    generator alone (random init + inference), and the generator's feed-forward output predictions."""
    gen_params = (generator, None)
    kw = dict(solver=ode_solver, max_t1=max_t1, dt=dt, stepsize_controller=stepsize_controller)
    amort_activities = init_activities_with_amort(amortiser, generator, output)
    amort_preds = amort_activities[0]
    hpc_preds = solve_inference(gen_params, amort_activities, output, **kw)[0][0]
    activities = init_activities_from_normal(key, layer_sizes, "unsupervised", batch_size, sigma, device=output.device)
    gen_preds = solve_inference(gen_params, activities, output, **kw)[0][0]
    output_preds = init_activities_with_ffwd(generator, input, skip_model=None)[-1]
    return (compute_accuracy(input, amort_preds), compute_accuracy(input, hpc_preds), compute_accuracy(input, gen_preds),
            output_preds)


test_hpc.__test__ = False
