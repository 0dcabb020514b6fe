"""optax-shaped gradient transformations (`init(params)`, `update(updates, state, params)`) for
the two optimisers the reference's examples use on this path: `optax.sgd` and `optax.adam`
(SURVEY Appendix B).  The update is elementwise host-framework work (torch `_foreach` kernels),
outside the timed kernels, exactly as the reference leaves it to optax (jpc/_train.py:234-242).
"""
from __future__ import annotations

from typing import NamedTuple

import torch

from .nn import tree_leaves, tree_unflatten_like


class GradientTransformation(NamedTuple):
    init: callable
    update: callable


def sgd(learning_rate: float) -> GradientTransformation:
    """optax.sgd(lr): updates = -lr * grads; stateless."""
    def init(params):
        return ()

    def update(updates, state, params=None):
        leaves = tree_leaves(updates)
        new = torch._foreach_mul(leaves, -float(learning_rate)) if leaves else []
        return tree_unflatten_like(updates, list(new)), state
    return GradientTransformation(init, update)


def adam(learning_rate: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> GradientTransformation:
    """optax.adam(lr): bias-corrected moments, eps added outside the sqrt (eps_root = 0)."""
    def init(params):
        leaves = tree_leaves(params)
        return {"count": 0, "mu": [torch.zeros_like(p) for p in leaves],
                "nu": [torch.zeros_like(p) for p in leaves]}

    def update(updates, state, params=None):
        g = tree_leaves(updates)
        count = state["count"] + 1
        mu = torch._foreach_mul(state["mu"], b1)
        torch._foreach_add_(mu, g, alpha=1 - b1)
        nu = torch._foreach_mul(state["nu"], b2)
        torch._foreach_addcmul_(nu, g, g, value=1 - b2)
        denom = torch._foreach_div(nu, 1 - b2 ** count)
        torch._foreach_sqrt_(denom)
        torch._foreach_add_(denom, eps)
        step = torch._foreach_div(mu, denom)
        torch._foreach_mul_(step, -float(learning_rate) / (1 - b1 ** count))
        return tree_unflatten_like(updates, list(step)), {"count": count, "mu": list(mu), "nu": list(nu)}
    return GradientTransformation(init, update)


def apply_updates(model, updates):
    """eqx.apply_updates: leafwise p + u, structure preserved."""
    p, u = tree_leaves(model), tree_leaves(updates)
    new = torch._foreach_add(p, u) if p else []
    return tree_unflatten_like(model, list(new))
