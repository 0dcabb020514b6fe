"""Error-reparameterised predictive coding (ePC, Goemaere et al. 2025) on the B200 path: `init_epc_errors`
(jpc/_core/_init.py:163-190), `epc_energy_fn` (jpc/_core/_energies.py:360-458), `compute_epc_error_grad` /
`compute_epc_param_grads` (jpc/_core/_grads.py:857-949), `update_epc_errors` / `update_epc_params`
(jpc/_core/_updates.py:785-960).  Every function is one `jpcb_epc_grads` call (forward chain + optional backward
chain on the device); supervised, or `x=None` with errors[0] in the role of z_0."""
from __future__ import annotations

import torch

from . import _lib
from ._core import _Net, _grads_pytree, _unpack_params
from .optim import update_and_apply


def init_epc_errors(layer_sizes, batch_size, mode="supervised", device=None):
    """jpc/_core/_init.py:163-190: zero errors for layers 1..L (plus layer 0 in "unsupervised" mode)."""
    if mode not in ("supervised", "unsupervised"):
        raise ValueError("`mode` must be 'supervised' or 'unsupervised'")
    device = torch.device("cuda" if device is None else device)
    start_l = 0 if mode == "unsupervised" else 1
    return [torch.zeros(int(batch_size), int(layer_sizes[l]), dtype=torch.float32, device=device)
            for l in range(start_l, len(layer_sizes))]


def _run(params, errors, y, x, loss, param_type, want_energy, want_egrads, want_pgrads):
    model, skip_model = _unpack_params(params)
    # x=None: errors[0] plays z_0 and the list has L + 1 entries (jpc/_core/_energies.py:429-432)
    net = _Net(model, skip_model, x, y, param_type=param_type, loss_id=loss, gamma=None, output_energy_scaling=None,
               activity_decay=0.0, weight_decay=0.0, spectral_penalty=0.0, scaled_free_input=True)
    net.check_acts(errors)  # the error list has the shapes of the activity list
    net.ws_bytes = _lib.lib.jpcb_epc_workspace_bytes(net.desc)
    ws = net.workspace()
    F = torch.empty((), dtype=torch.float32, device=net.dev) if want_energy else None
    gE = net.new_acts() if want_egrads else None
    gW = [torch.empty(s, dtype=torch.float32, device=net.dev) for s in net.w_shapes] if want_pgrads else None
    gb = [torch.empty(s, dtype=torch.float32, device=net.dev) for s in net.b_shapes] if want_pgrads and net.has_bias else None
    _lib.check(_lib.lib.jpcb_epc_grads(net.desc, net.W, net.b, _lib.ptr_array(net.in_acts(errors)), _lib.ptr(net.x), _lib.ptr(net.y),
                                       _lib.ptr(F), _lib.ptr_array(gE) if gE else None, _lib.ptr_array(gW) if gW else None,
                                       _lib.ptr_array(gb) if gb else None, _lib.ptr(ws), ws.numel(), net.stream()))
    egrads = net.out_acts(gE) if want_egrads else None
    pgrads = _grads_pytree(net, *net.out_grads(gW, gb)) if want_pgrads else None
    return F, egrads, pgrads


def epc_energy_fn(params, errors, y, *, x=None, loss="mse", param_type="sp"):
    """jpc/_core/_energies.py:360-458: total ePC energy / batch size."""
    return _run(params, errors, y, x, loss, param_type, True, False, False)[0]


def compute_epc_error_grad(params, errors, y, *, x=None, loss_id="mse", param_type="sp"):
    """jpc/_core/_grads.py:857-904: (energy, dF/d errors)."""
    F, g, _ = _run(params, errors, y, x, loss_id, param_type, True, True, False)
    return F, g


def compute_epc_param_grads(params, errors, y, *, x=None, loss_id="mse", param_type="sp"):
    """jpc/_core/_grads.py:907-949: (model_grads, skip_grads) with the models' structure."""
    return _run(params, errors, y, x, loss_id, param_type, False, False, True)[2]


def update_epc_errors(params, errors, optim, opt_state, output, *, input=None, loss_id="mse", param_type="sp"):
    """jpc/_core/_updates.py:785-877."""
    energy, grads = compute_epc_error_grad(params, errors, output, x=input, loss_id=loss_id, param_type=param_type)
    errors, opt_state = update_and_apply(optim, grads, opt_state, errors)
    return {"energy": energy, "errors": errors, "grads": grads, "opt_state": opt_state}


def update_epc_params(params, errors, optim, opt_state, output, *, input=None, loss_id="mse", param_type="sp"):
    """jpc/_core/_updates.py:880-960."""
    grads = compute_epc_param_grads(params, errors, output, x=input, loss_id=loss_id, param_type=param_type)
    (model, skip_model), opt_state = update_and_apply(optim, grads, opt_state, tuple(params))
    return {"model": model, "skip_model": skip_model, "grads": grads, "opt_state": opt_state}
