"""User-side helpers of jpc/_utils.py:129-260 (losses, accuracy, recorded-trajectory helpers, norms).  These are
evaluation/metric utilities outside the train-step kernels; they run as torch device ops on
whatever device their inputs live on."""
from __future__ import annotations

import torch

from .nn import tree_leaves


def mse_loss(preds, labels):
    """jpc/_utils.py:129-130."""
    return 0.5 * torch.mean(torch.sum((labels - preds) ** 2, dim=1))


def cross_entropy_loss(logits, labels):
    """jpc/_utils.py:133-136."""
    return -torch.mean(torch.sum(labels * torch.log(torch.softmax(logits, dim=-1)), dim=-1))


def compute_accuracy(truths, preds):
    """jpc/_utils.py:139-142."""
    return torch.mean((torch.argmax(truths, dim=1) == torch.argmax(preds, dim=1)).float()) * 100


def compute_activity_norms(activities):
    """jpc/_utils.py:219-230: mean over the batch of the per-sample l2 norm, per layer."""
    return torch.stack([torch.linalg.vector_norm(a, ord=2, dim=-1).mean() for a in tree_leaves(activities)])


def compute_param_norms(params):
    """jpc/_utils.py:233-260: l2 norm of every array leaf of (model, skip_model)."""
    def proc(tree):
        leaves = tree_leaves(tree)
        if not leaves:
            return torch.zeros(0)
        return torch.stack([torch.linalg.vector_norm(p.reshape(-1), ord=2) for p in leaves])
    model_params, skip_model_params = params
    return proc(model_params), (proc(skip_model_params) if skip_model_params is not None else None)


def get_t_max(activities_iters):
    """jpc/_utils.py:145-146: index of the last recorded step = (first inf row of the first leaf) - 1.
    Stays on the device like the reference's traced value; rows past the last step are inf
    (`solve_inference(record_iters=True)`)."""
    return torch.argmax(activities_iters[0][:, 0, 0]) - 1


def compute_infer_energies(params, activities_iters, t_max, y, *, x=None, loss="mse", param_type="sp", weight_decay=0.0,
                           spectral_penalty=0.0, activity_decay=0.0):
    """jpc/_utils.py:149-216: layer energies at recorded steps 0..t_max-1, as a (len(model), 500) array that is
    zero past t_max, rows in the reference's order (`pc_energy_fn(record_layers=True)` reversed, `:216`).
    One `jpcb_pc_energy` call per recorded step."""
    from ._core import pc_energy_fn
    model, _ = params
    out = torch.zeros(len(model), 500, dtype=torch.float32, device=y.device)
    for t in range(min(int(t_max), 500)):
        out[:, t] = pc_energy_fn(params, [a[t] for a in activities_iters], y, x=x, loss=loss, param_type=param_type,
                                 weight_decay=weight_decay, spectral_penalty=spectral_penalty,
                                 activity_decay=activity_decay, record_layers=True)
    return out.flip(0)
