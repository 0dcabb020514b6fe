"""User-side helpers of jpc/_utils.py:129-142,219-260 (losses, accuracy, norms).  These are
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
