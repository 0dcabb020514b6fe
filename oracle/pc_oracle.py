"""CPU oracle for the predictive-coding (PC) train-step hot path of thebuckleylab/jpc.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
(`jpc_b200`) never imports, links or falls back to anything in this directory.

What it is: a torch (CPU) restatement of the reference's algorithm for the hot path.  The
reference differentiates a Python energy function with ``jax.grad``; this oracle restates that
ENERGY (same terms, same order, same quirks) and differentiates it with ``torch.autograd``, so it
has the same structure as the reference.  A second, independent closed-form implementation of
the gradients is provided and the tests require both to agree to ~1e-13 in
fp64.  dtype follows the inputs: fp64 is the arbiter, fp32 shows the noise floor.

PARITY PINNING STATUS ("parity unpinned" for most of the path):
  * JAX / equinox / diffrax / optax are not installable in this image (no network), so the
    reference cannot be executed to generate golden vectors, and the reference's own tests hold
    no numeric expectations for this path except three analytical identities on deep linear
    networks (reference ``tests/test_analytical.py:162-224,227-290,293-368``).
  * Pinned: those known answers are restated here (`linear_equilib_energy`,
    `linear_activity_solution`) following ``jpc/_core/_analytical.py:15-69,229-269`` and the
    oracle's Euler inference reproduces them (tests/test_oracle_kat.py) at the reference's own
    tolerance (1e-2) and far tighter (1e-6).
  * Unpinned: diffrax Euler/Heun/ConstantStepSize step semantics (third-party, restated from
    upstream diffrax 0.6-0.7 behaviour), non-linear activations, biases, skips, bPC.

Third-party semantics restated (not under /root/reference): diffrax>=0.6.0 (pyproject.toml:38)
``Euler``/``Heun``/``ConstantStepSize``/``Event``; optax>=0.2.4 ``sgd``/``adam``;
equinox>=0.11.2 ``nn.Linear`` (y = W x + b, weight[out,in]); optimistix ``rms_norm``.

Model representation (plain data, independent of jpc_b200): a "model" is a list of layer dicts
``{"W": [out,in], "b": [out] or None, "act": str}`` in the reference's ``make_mlp`` form
f_l(u) = W_l act_l(u) + b_l (``jpc/_utils.py:87-106``), ``act`` of layer 0 being "linear".
``skips`` is a list of bools (identity skip at layer l) or None (``jpc/_utils.py:122-126``).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

PARAM_TYPES = ("sp", "mupc", "ntp")  # jpc/_core/_errors.py:1
ACT_FNS = ("linear", "tanh", "hard_tanh", "relu", "leaky_relu", "gelu", "selu", "silu")  # jpc/_utils.py:13-15

_SELU_ALPHA = 1.6732632423543772848170429916717
_SELU_SCALE = 1.0507009873554804934193349852946


def check_param_type(param_type: str) -> None:
    """jpc/_core/_errors.py:4-11."""
    if param_type not in PARAM_TYPES:
        raise ValueError(f"Invalid param_type {param_type!r}; options are {PARAM_TYPES}")


def act_fn(name: str, x: torch.Tensor) -> torch.Tensor:
    """The 8 activation ids of jpc/_utils.py:18-38 with jax.nn conventions."""
    if name == "linear":
        return x
    if name == "tanh":
        return torch.tanh(x)
    if name == "hard_tanh":  # jax.nn.hard_tanh = where(x>1, 1, where(x<-1, -1, x))
        return torch.where(x > 1, torch.ones_like(x), torch.where(x < -1, -torch.ones_like(x), x))
    if name == "relu":  # derivative 0 at 0 in both jax and torch
        return torch.relu(x)
    if name == "leaky_relu":  # jax: where(x >= 0, x, 0.01 x)
        return torch.where(x >= 0, x, 0.01 * x)
    if name == "gelu":  # jax.nn.gelu default approximate=True (tanh form)
        return F.gelu(x, approximate="tanh")
    if name == "selu":
        safe = torch.where(x > 0, torch.zeros_like(x), x)
        return _SELU_SCALE * torch.where(x > 0, x, _SELU_ALPHA * torch.expm1(safe))
    if name == "silu":
        return x * torch.sigmoid(x)
    raise ValueError(f"Invalid activation function ID. Options are {ACT_FNS}.")


def act_deriv(name: str, x: torch.Tensor) -> torch.Tensor:
    """Closed-form derivative of `act_fn` (used only by the closed-form cross-check)."""
    one, zero = torch.ones_like(x), torch.zeros_like(x)
    if name == "linear":
        return one
    if name == "tanh":
        return 1 - torch.tanh(x) ** 2
    if name == "hard_tanh":
        return torch.where((x > 1) | (x < -1), zero, one)
    if name == "relu":
        return torch.where(x > 0, one, zero)
    if name == "leaky_relu":
        return torch.where(x >= 0, one, 0.01 * one)
    if name == "gelu":
        c = math.sqrt(2.0 / math.pi)
        u = c * (x + 0.044715 * x ** 3)
        t = torch.tanh(u)
        return 0.5 * (1 + t) + 0.5 * x * (1 - t * t) * c * (1 + 3 * 0.044715 * x * x)
    if name == "selu":
        return _SELU_SCALE * torch.where(x > 0, one, _SELU_ALPHA * torch.exp(x))
    if name == "silu":
        s = torch.sigmoid(x)
        return s * (1 + x * (1 - s))
    raise ValueError(name)


# ----------------------------------------------------------------------------------------------
# model helpers
# ----------------------------------------------------------------------------------------------
def make_mlp(seed: int, input_dim: int, width: int, depth: int, output_dim: int, act: str,
             use_bias: bool = False, param_type: str = "sp", dtype=torch.float64) -> List[dict]:
    """Synthetic weights with the law of jpc/_utils.py:87-106 (equinox Linear default
    U(-1/sqrt(in), 1/sqrt(in)); N(0,1) weights for mupc).  The RNG stream is torch's, not JAX's."""
    check_param_type(param_type)
    if act not in ACT_FNS:
        raise ValueError(f"Invalid activation function ID. Options are {ACT_FNS}.")
    g = torch.Generator().manual_seed(seed)
    layers = []
    for i in range(depth):
        _in = input_dim if i == 0 else width
        _out = output_dim if i + 1 == depth else width
        lim = 1.0 / math.sqrt(_in)
        W = (torch.rand(_out, _in, generator=g, dtype=torch.float64) * 2 - 1) * lim
        b = (torch.rand(_out, generator=g, dtype=torch.float64) * 2 - 1) * lim if use_bias else None
        if param_type == "mupc":
            W = torch.randn(_out, _in, generator=g, dtype=torch.float64)
        layers.append({"W": W.to(dtype), "b": None if b is None else b.to(dtype),
                       "act": "linear" if i == 0 else act})
    return layers


def make_skip_model(depth: int) -> List[bool]:
    """jpc/_utils.py:111-126: identity skips at layers 1..depth-2."""
    return [0 < l < depth - 1 for l in range(depth)]


def layer_fwd(layer: dict, u: torch.Tensor) -> torch.Tensor:
    """vmap(model[l])(u) for the make_mlp layer form: W act(u) + b (jpc/_utils.py:102-106)."""
    out = act_fn(layer["act"], u) @ layer["W"].T
    if layer["b"] is not None:
        out = out + layer["b"]
    return out


def get_param_scalings(model: Sequence[dict], x: Optional[torch.Tensor], skips=None,
                       param_type: str = "sp", gamma: Optional[float] = None) -> List[float]:
    """jpc/_core/_energies.py:879-933."""
    L = len(model)
    no_skips = skips is None or not any(skips)
    if param_type == "sp":
        scalings = [1.0] * L
    else:
        D = x.shape[1]
        N = model[0]["W"].shape[0]
        a1 = 1 / math.sqrt(D)
        al = 1 / math.sqrt(N) if no_skips else 1 / math.sqrt(N * L)
        aL = 1 / N if param_type == "mupc" else 1 / math.sqrt(N)
        scalings = [a1] + [al] * (L - 2) + [aL]
    if gamma is not None:
        scalings[-1] = scalings[-1] / gamma
    return scalings


# ----------------------------------------------------------------------------------------------
# losses (jpc/_utils.py:129-136)
# ----------------------------------------------------------------------------------------------
def mse_loss(preds, labels):
    return 0.5 * torch.mean(torch.sum((labels - preds) ** 2, dim=1))


def cross_entropy_loss(logits, labels):
    return -torch.mean(torch.sum(labels * torch.log(torch.softmax(logits, dim=-1)), dim=-1))


# ----------------------------------------------------------------------------------------------
# PC energy  (jpc/_core/_energies.py:13-158)
# ----------------------------------------------------------------------------------------------
def pc_energy_fn(model, skips, activities, y, *, x=None, loss="mse", param_type="sp",
                 activity_decay=0.0, record_layers=False, gamma=None, output_energy_scaling=None):
    """Total (or per-layer) PC energy normalised by batch size.

    weight_decay / spectral_penalty are omitted on purpose: for make_mlp models they are no-ops
    in the reference (layers are nn.Sequential, which has no `.weight`; _energies.py:128-143).
    """
    check_param_type(param_type)
    batch_size = y.shape[0]
    start_activity_l = 1 if x is not None else 2
    n_activity_layers = len(activities) - 1
    n_hidden = len(model) - 1
    if skips is None:
        skips = [False] * len(model)
    scalings = get_param_scalings(model, x, skips, param_type, gamma)
    output_scale = 1.0 if output_energy_scaling is None else output_energy_scaling

    if loss == "mse":
        eL = y - scalings[-1] * layer_fwd(model[-1], activities[-2])
        energies = [0.5 * output_scale * torch.sum(eL ** 2)]
    elif loss == "ce":
        logits = scalings[-1] * layer_fwd(model[-1], activities[-2])
        energies = [-output_scale * torch.sum(y * torch.log_softmax(logits, dim=-1))]
    else:
        raise ValueError("`mse` and `ce` are the only valid losses.")

    for act_l, net_l in zip(range(start_activity_l, n_activity_layers), range(1, n_hidden)):
        err = activities[act_l] - scalings[net_l] * layer_fwd(model[net_l], activities[act_l - 1])
        if skips[net_l]:
            err = err - activities[act_l - 1]
        energies.append(0.5 * torch.sum(err ** 2))

    if x is not None:
        e1 = activities[0] - scalings[0] * layer_fwd(model[0], x)
    else:  # unscaled first layer in the unsupervised branch (_energies.py:123-124)
        e1 = activities[1] - layer_fwd(model[0], activities[0])
    energies.append(0.5 * torch.sum(e1 ** 2))

    activity_reg = 0.0
    if activity_decay > 0.0:  # over ALL entries incl. the dummy output slot (_energies.py:145-150)
        for a in activities:
            activity_reg = activity_reg + torch.sum(a ** 2, dim=1).mean()
        activity_reg = activity_reg * activity_decay / 2

    stacked = torch.stack(energies)
    if record_layers:
        return stacked / batch_size
    return stacked.sum() / batch_size + activity_reg


def init_activities_with_ffwd(model, x, *, skips=None, param_type="sp", gamma=None):
    """jpc/_core/_init.py:56-82."""
    check_param_type(param_type)
    L = len(model)
    if skips is None:
        skips = [False] * L
    scalings = get_param_scalings(model, x, skips, param_type, gamma)
    z = scalings[0] * layer_fwd(model[0], x)
    if skips[0]:
        z = z + x
    acts = [z]
    for l in range(1, L):
        z = scalings[l] * layer_fwd(model[l], acts[l - 1])
        if skips[l]:
            z = z + acts[l - 1]
        acts.append(z)
    return acts


def compute_pc_activity_grad(model, skips, activities, y, **kw):
    """value_and_grad(pc_energy_fn, argnums=1)  (jpc/_core/_grads.py:154-167)."""
    acts = [a.detach().clone().requires_grad_(True) for a in activities]
    energy = pc_energy_fn(model, skips, acts, y, **kw)
    grads = torch.autograd.grad(energy, acts, allow_unused=True)
    grads = [torch.zeros_like(a) if g is None else g for a, g in zip(acts, grads)]
    return energy.detach(), grads


def compute_pc_param_grads(model, skips, activities, y, **kw):
    """filter_grad(pc_energy_fn) wrt the model leaves (jpc/_core/_grads.py:487-499)."""
    m2, leaves = [], []
    for layer in model:
        W = layer["W"].detach().clone().requires_grad_(True)
        b = None if layer["b"] is None else layer["b"].detach().clone().requires_grad_(True)
        m2.append({"W": W, "b": b, "act": layer["act"]})
        leaves.append(W)
        if b is not None:
            leaves.append(b)
    energy = pc_energy_fn(m2, skips, [a.detach() for a in activities], y, **kw)
    g = list(torch.autograd.grad(energy, leaves, allow_unused=True))
    out = []
    for layer in m2:
        gW = g.pop(0)
        gb = g.pop(0) if layer["b"] is not None else None
        out.append({"W": torch.zeros_like(layer["W"]) if gW is None else gW,
                    "b": gb})
    return out


# ----------------------------------------------------------------------------------------------
# closed forms -- independent cross-check of the autograd path, mse only
# ----------------------------------------------------------------------------------------------
def closed_form_errors(model, skips, activities, y, x, param_type="sp", gamma=None,
                       output_energy_scaling=None):
    """e_{l+1} = target_{l+1} - a_l (W_l act_l(u_l) + b_l) - s_l u_l for l=0..L-1 (u_0=x, u_l=z_l)."""
    L = len(model)
    if skips is None:
        skips = [False] * L
    a = get_param_scalings(model, x, skips, param_type, gamma)
    errs = []
    for l in range(L):
        u = x if l == 0 else activities[l - 1]
        tgt = y if l == L - 1 else activities[l]
        e = tgt - a[l] * layer_fwd(model[l], u)
        if skips[l] and 1 <= l <= L - 2:
            e = e - u
        errs.append(e)
    return errs, a


def closed_form_activity_grad(model, skips, activities, y, x, param_type="sp", gamma=None,
                              output_energy_scaling=None, activity_decay=0.0):
    L = len(model)
    B = y.shape[0]
    lam = 1.0 if output_energy_scaling is None else output_energy_scaling
    if skips is None:
        skips = [False] * L
    errs, a = closed_form_errors(model, skips, activities, y, x, param_type, gamma)
    errs[-1] = lam * errs[-1]
    grads = []
    for l in range(1, L):  # z_l feeds layer l
        z = activities[l - 1]
        back = (errs[l] @ model[l]["W"]) * a[l] * act_deriv(model[l]["act"], z)
        g = errs[l - 1] - back
        if skips[l] and 1 <= l <= L - 2:
            g = g - errs[l]
        grads.append(g / B + activity_decay * z / B)
    grads.append(activity_decay * activities[-1] / B)  # dummy slot
    return grads


def closed_form_param_grads(model, skips, activities, y, x, param_type="sp", gamma=None,
                            output_energy_scaling=None):
    L = len(model)
    B = y.shape[0]
    lam = 1.0 if output_energy_scaling is None else output_energy_scaling
    errs, a = closed_form_errors(model, skips, activities, y, x, param_type, gamma)
    errs[-1] = lam * errs[-1]
    out = []
    for l in range(L):
        u = x if l == 0 else activities[l - 1]
        h = act_fn(model[l]["act"], u)
        gW = -(a[l] / B) * errs[l].T @ h
        gb = None if model[l]["b"] is None else -(a[l] / B) * errs[l].sum(0)
        out.append({"W": gW, "b": gb})
    return out


# ----------------------------------------------------------------------------------------------
# fixed-step inference  (jpc/_core/_infer.py:103-145 + diffrax semantics)
# ----------------------------------------------------------------------------------------------
def n_fixed_steps(max_t1: float, dt: float) -> int:
    """ConstantStepSize: steps of dt from t0=0 while t < t1, the last one clipped to t1."""
    return max(1, int(math.ceil(max_t1 / dt - 1e-9)))


def rms_norm(tensors) -> float:
    """optimistix.rms_norm over a pytree: sqrt(sum x^2 / total size)."""
    tot = sum(float((t.double() ** 2).sum()) for t in tensors)
    n = sum(t.numel() for t in tensors)
    return math.sqrt(tot / n)


def solve_inference(model, skips, activities, y, *, x, solver="euler", max_t1=20.0, dt=0.5,
                    event_atol=1e-3, event_rtol=1e-3, use_closed_form=False, **kw):
    """Fixed-step (ConstantStepSize) Euler or Heun solve of dz/dt = -grad_z F from t=0 to max_t1.

    Returns (activities_at_t1, n_steps_taken).  The steady-state Event of the reference
    (`rms_norm(y) < atol + rtol*rms_norm(y)`, _infer.py:136-145) is honoured after each step.
    """
    acts = [a.detach().clone() for a in activities]

    def vf(a_list):
        if use_closed_form:
            g = closed_form_activity_grad(model, skips, a_list, y, x, **kw)
        else:
            _, g = compute_pc_activity_grad(model, skips, a_list, y, x=x, **kw)
        return [-gi for gi in g]

    t, steps = 0.0, 0
    while t < max_t1 - 1e-12:
        h = min(dt, max_t1 - t)
        k1 = vf(acts)
        if solver == "euler":
            acts = [a + h * k for a, k in zip(acts, k1)]
        elif solver == "heun":
            tmp = [a + h * k for a, k in zip(acts, k1)]
            k2 = vf(tmp)
            acts = [a + h * (0.5 * ka + 0.5 * kb) for a, ka, kb in zip(acts, k1, k2)]
        else:
            raise ValueError(solver)
        t += h
        steps += 1
        r = rms_norm(acts)
        if r < event_atol + event_rtol * r:
            break
    return acts, steps


def solve_inference_adaptive(model, skips, activities, y, *, x, max_t1=20.0, dt0=None, rtol=1e-3, atol=1e-3,
                             safety=0.9, factormin=0.2, factormax=10.0, max_steps=4096, **kw):
    """diffrax.Heun + PIDController(rtol, atol) (I-controller defaults) + the reference's steady-state Event:
    what `jpc.solve_inference` does when called with its defaults (jpc/_core/_infer.py:28-33,109-132).

    diffrax is a third-party dependency that is not under /root/reference; the controller below restates
    upstream diffrax 0.6-0.7 from memory: embedded error y_err = dt (k1 - k2) / 2, scaled
    error = rms_norm(y_err / (atol + rtol max(|y0|, |y1|))), accept iff < 1, factor = clip(safety *
    err^(-1/order), [1 if accepted else factormin, factormax]) with order = 2; Hairer initial step when dt0 is
    None.  PARITY UNPINNED against a live diffrax run.  Returns (activities, stats)."""
    acts = [a.detach().clone() for a in activities]
    order = 2.0

    def vf(a_list):
        _, g = compute_pc_activity_grad(model, skips, a_list, y, x=x, **kw)
        return [-gi for gi in g]

    if dt0 is None:
        f0 = vf(acts)
        scale = [atol + a.abs() * rtol for a in acts]
        d0 = rms_norm([a / s for a, s in zip(acts, scale)])
        d1 = rms_norm([f / s for f, s in zip(f0, scale)])
        small = d0 < 1e-5 or d1 < 1e-5
        h0 = 1e-6 if small else 0.01 * d0 / d1
        f1 = vf([a + h0 * f for a, f in zip(acts, f0)])
        d2 = rms_norm([(fb - fa) / s for fa, fb, s in zip(f0, f1, scale)]) / h0
        max_d = max(d1, d2)
        h1 = max(1e-6, h0 * 1e-3) if max_d <= 1e-15 else (0.01 / max_d) ** (1.0 / order)
        dt = min(100.0 * h0, h1)
    else:
        dt = float(dt0)
    t, steps = 0.0, 0
    stats = {"accepted": 0, "rejected": 0, "event": False, "dts": []}
    while t < max_t1:
        if steps >= max_steps:
            raise RuntimeError("The maximum number of solver steps was reached. Try increasing `max_steps`.")
        h = min(dt, max_t1 - t)
        k1 = vf(acts)
        k2 = vf([a + h * k for a, k in zip(acts, k1)])
        cand = [a + h * (0.5 * ka + 0.5 * kb) for a, ka, kb in zip(acts, k1, k2)]
        yerr = [h * (0.5 * ka - 0.5 * kb) for ka, kb in zip(k1, k2)]
        err = rms_norm([e / (atol + rtol * torch.maximum(a.abs(), c.abs())) for e, a, c in zip(yerr, acts, cand)])
        keep = err < 1.0
        inv = float("inf") if err == 0.0 else 1.0 / err
        factor = min(max(safety * inv ** (1.0 / order), 1.0 if keep else factormin), factormax)
        dt = h * factor
        steps += 1
        if keep:
            t = max_t1 if max_t1 - (t + h) < 1e-6 * max(1.0, max_t1) else t + h
            acts = cand
            stats["accepted"] += 1
            stats["dts"].append(h)
            r = rms_norm(acts)
            if r < atol + rtol * r:
                stats["event"] = True
                break
        else:
            stats["rejected"] += 1
    return acts, stats


def pc_train_step(model, skips, x, y, *, solver="euler", max_t1=20.0, dt=0.5, loss_id="mse",
                  param_type="sp", activity_decay=0.0):
    """What make_pc_step computes before the optimiser (jpc/_train.py:135-232): ffwd init, loss at
    init, inference, per-layer energies at the solution, parameter gradients."""
    acts0 = init_activities_with_ffwd(model, x, skips=skips, param_type=param_type)
    if loss_id == "mse":
        loss = mse_loss(acts0[-1], y)
    elif loss_id == "ce":
        loss = cross_entropy_loss(acts0[-1], y)
    else:
        raise ValueError("`mse` and `ce` are the only valid losses.")
    acts, steps = solve_inference(model, skips, acts0, y, x=x, solver=solver, max_t1=max_t1, dt=dt,
                                  loss=loss_id, param_type=param_type, activity_decay=activity_decay)
    energies = pc_energy_fn(model, skips, acts, y, x=x, loss=loss_id, param_type=param_type,
                            activity_decay=activity_decay, record_layers=True)
    grads = compute_pc_param_grads(model, skips, acts, y, x=x, loss=loss_id, param_type=param_type,
                                   activity_decay=activity_decay)
    return {"loss": loss, "activities": acts, "energies": energies.detach(), "grads": grads,
            "init_activities": acts0, "n_steps": steps}


# ----------------------------------------------------------------------------------------------
# bidirectional PC  (jpc/_core/_energies.py:246-357)
# ----------------------------------------------------------------------------------------------
def bpc_energy_fn(top_down, bottom_up, activities, y, *, x=None, skips=None, param_type="sp",
                  record_layers=False, backward_energy_weight=1.0, forward_energy_weight=1.0):
    check_param_type(param_type)  # validated but no scalings are applied (_energies.py:314)
    batch_size = activities[0].shape[0]
    H = len(top_down) - 1
    if skips is None:
        skips = [False] * (H + 1)
    x_eff = activities[0] if x is None else x
    pairs = []
    delta_0 = x_eff - layer_fwd(bottom_up[0], activities[0])
    e_L = y - layer_fwd(top_down[-1], activities[-2])
    pairs.append((forward_energy_weight * 0.5 * torch.sum(e_L ** 2),
                  backward_energy_weight * 0.5 * torch.sum(delta_0 ** 2)))
    for l in range(H):
        act_prev = x_eff if l == 0 else activities[l - 1]
        e_l = activities[l] - layer_fwd(top_down[l], act_prev)
        if skips[l] and l > 0:
            e_l = e_l - act_prev
        act_next = y if l == H - 1 else activities[l + 1]
        d_l = activities[l] - layer_fwd(bottom_up[l + 1], act_next)
        if skips[l + 1] and l < H - 1:
            d_l = d_l - act_next
        pairs.append((forward_energy_weight * 0.5 * torch.sum(e_l ** 2),
                      backward_energy_weight * 0.5 * torch.sum(d_l ** 2)))
    flat = torch.stack([v for p in pairs for v in p])
    if record_layers:
        return flat / batch_size
    return flat.sum() / batch_size


def compute_bpc_activity_grad(top_down, bottom_up, activities, y, **kw):
    """jpc/_core/_grads.py:215-226."""
    acts = [a.detach().clone().requires_grad_(True) for a in activities]
    energy = bpc_energy_fn(top_down, bottom_up, acts, y, **kw)
    grads = torch.autograd.grad(energy, acts, allow_unused=True)
    grads = [torch.zeros_like(a) if g is None else g for a, g in zip(acts, grads)]
    return energy.detach(), grads


def compute_bpc_param_grads(top_down, bottom_up, activities, y, **kw):
    """jpc/_core/_grads.py:589-612."""
    def clone(model):
        m2, leaves = [], []
        for layer in model:
            W = layer["W"].detach().clone().requires_grad_(True)
            b = None if layer["b"] is None else layer["b"].detach().clone().requires_grad_(True)
            m2.append({"W": W, "b": b, "act": layer["act"]})
            leaves.append(W)
            if b is not None:
                leaves.append(b)
        return m2, leaves
    td, td_leaves = clone(top_down)
    bu, bu_leaves = clone(bottom_up)
    energy = bpc_energy_fn(td, bu, [a.detach() for a in activities], y, **kw)
    g = list(torch.autograd.grad(energy, td_leaves + bu_leaves, allow_unused=True))

    def unpack(m2):
        out = []
        for layer in m2:
            gW = g.pop(0)
            gb = g.pop(0) if layer["b"] is not None else None
            out.append({"W": torch.zeros_like(layer["W"]) if gW is None else gW, "b": gb})
        return out
    return unpack(td), unpack(bu)


# ----------------------------------------------------------------------------------------------
# optimisers (optax semantics) -- for the discrete update twins
# ----------------------------------------------------------------------------------------------
def sgd_update(params, grads, lr):
    """optax.sgd(lr): update = -lr * g; eqx.apply_updates: p + update."""
    return [p - lr * g for p, g in zip(params, grads)]


def adam_update(params, grads, state, lr, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam(lr): bias-corrected, eps outside the sqrt (eps_root=0)."""
    if state is None:
        state = {"count": 0, "mu": [torch.zeros_like(p) for p in params],
                 "nu": [torch.zeros_like(p) for p in params]}
    count = state["count"] + 1
    mu = [b1 * m + (1 - b1) * g for m, g in zip(state["mu"], grads)]
    nu = [b2 * v + (1 - b2) * g * g for v, g in zip(state["nu"], grads)]
    new = []
    for p, m, v in zip(params, mu, nu):
        mhat = m / (1 - b1 ** count)
        vhat = v / (1 - b2 ** count)
        new.append(p - lr * mhat / (torch.sqrt(vhat) + eps))
    return new, {"count": count, "mu": mu, "nu": nu}


# ----------------------------------------------------------------------------------------------
# analytical known answers for deep LINEAR networks (jpc/_core/_analytical.py)
# ----------------------------------------------------------------------------------------------
def linear_equilib_rescaling(Ws, scalings, output_energy_scaling=1.0, use_skips=False):
    """S = I/lambda + sum_l P_l P_l^T  (jpc/_core/_analytical.py:15-69)."""
    L = len(Ws)
    d_y = Ws[-1].shape[0]
    width = Ws[0].shape[0]
    eye = lambda n: torch.eye(n, dtype=Ws[0].dtype)
    S = eye(d_y) / output_energy_scaling
    if use_skips:
        S = S + (scalings[-1] ** 2) * (Ws[-1] @ Ws[-1].T)
        for l in range(1, L - 1):
            prod = eye(width)
            for k in range(l, L - 1):
                prod = (eye(width) + scalings[k] * Ws[k]) @ prod
            p = scalings[-1] * Ws[-1] @ prod
            S = S + p @ p.T
    else:
        cum = eye(d_y)
        for i in range(L - 1, 0, -1):
            cum = cum @ Ws[i]
            cs = 1.0
            for j in range(i, L):
                cs *= scalings[j]
            S = S + (cs ** 2) * (cum @ cum.T)
    return S


def linear_equilib_energy(model, skips, x, y, *, param_type="sp", gamma=None,
                          output_energy_scaling=None):
    """F* = mean_i 1/2 r_i^T S^{-1} r_i, r = y - W_{L:1} x  (jpc/_core/_analytical.py:229-269)."""
    check_param_type(param_type)
    scalings = get_param_scalings(model, x, skips, param_type, gamma)
    Ws = [layer["W"] for layer in model]
    L = len(Ws)
    N = Ws[0].shape[0]
    has_skips = skips is not None and any(skips)
    lam = 1.0 if output_energy_scaling is None else output_energy_scaling
    eye = lambda n: torch.eye(n, dtype=Ws[0].dtype)
    if has_skips:
        W1 = scalings[0] * Ws[0]
        hp = eye(N)
        for l in range(1, L - 1):
            hp = (eye(N) + scalings[l] * Ws[l]) @ hp
        WLto1 = scalings[-1] * Ws[-1] @ hp @ W1
    else:
        sp = 1.0
        for s in scalings:
            sp *= s
        WLto1 = eye(Ws[-1].shape[0])
        for i in range(L - 1, -1, -1):
            WLto1 = WLto1 @ Ws[i]
        WLto1 = sp * WLto1
    S = linear_equilib_rescaling(Ws, scalings, lam, has_skips)
    r = y - x @ WLto1.T
    sol = torch.linalg.solve(S, r.T).T
    return (0.5 * (r * sol).sum(1)).mean()


def linear_activity_solution(model, skips, x, y, **kw):
    """Exact minimiser of the (quadratic) energy of a linear network in the hidden activities,
    obtained by solving grad=0 with the autograd Hessian.  Same known answer as
    compute_linear_activity_solution (jpc/_core/_analytical.py:551-697), derived independently."""
    acts0 = init_activities_with_ffwd(model, x, skips=skips, param_type=kw.get("param_type", "sp"),
                                      gamma=kw.get("gamma"))
    hidden = acts0[:-1]
    sizes = [h.shape[1] for h in hidden]
    B = x.shape[0]
    sols = []
    for i in range(B):
        xi, yi = x[i:i + 1], y[i:i + 1]

        def fe(flat):
            parts = list(torch.split(flat, sizes))
            acts = [p[None, :] for p in parts] + [acts0[-1][i:i + 1]]
            return pc_energy_fn(model, skips, acts, yi, x=xi, **kw)
        z0 = torch.zeros(sum(sizes), dtype=x.dtype)
        Hm = torch.autograd.functional.hessian(fe, z0)
        g0 = torch.autograd.functional.jacobian(fe, z0)
        sols.append(torch.linalg.solve(Hm, -g0))
    Z = torch.stack(sols)
    out = list(torch.split(Z, sizes, dim=1))
    scal = get_param_scalings(model, x, skips, kw.get("param_type", "sp"), kw.get("gamma"))
    out.append(scal[-1] * layer_fwd(model[-1], out[-1]))
    return out


# ----------------------------------------------------------------------------------------------
# error metric used by every parity test (SURVEY 8d "parity gates")
# ----------------------------------------------------------------------------------------------
def max_rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b|; exact zeros in the oracle must be (near) exact zeros in `a`."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = float(b.abs().max())
    diff = float((a - b).abs().max())
    if denom == 0.0:
        return 0.0 if diff <= 1e-12 else float("inf")
    return diff / denom
