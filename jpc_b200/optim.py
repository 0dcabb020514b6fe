"""optax-shaped gradient transformations (`init(params)`, `update(updates, state, params)`) for
the two optimisers the reference's examples use on this path: `optax.sgd` and `optax.adam`
(SURVEY Appendix B).  Two ways in:
  * the optax protocol (`update` then `apply_updates`), which works for any tree of tensors on any
    device (activities too: `update_pc_activities` steps them with `optax.sgd`) and runs as torch
    `_foreach` ops -- the reference leaves exactly this to optax (jpc/_train.py:234-242);
  * `fused_apply(params, grads, state)`, used by `make_pc_step` / `update_pc_params` when the parameters
    are fp32 CUDA tensors: `optim.update` + `eqx.apply_updates` as ONE pass of the library's optimiser
    kernel (`jpcb_optim_sgd` / `jpcb_optim_adam`, include/jpc_b200.h), no intermediate `updates` tree.
Both give the same numbers (tests/test_gpu_parity.py::test_fused_optimiser_matches_optax_protocol).
"""
from __future__ import annotations

from typing import NamedTuple

import torch

from . import _flat
from .nn import tree_leaves, tree_unflatten_like


class GradientTransformation(NamedTuple):
    init: callable
    update: callable
    fused_apply: callable = None  # (params_tree, grads_tree, state) -> (new_params_tree, new_state), or None
    fused_apply_flat: callable = None  # the same on flat leaf lists: (p_leaves, g_leaves, state) -> (new_leaves, new_state) | None
    sgd_learning_rate: float = None    # set by sgd(): lets PCTrainer put the (stateless) update inside its CUDA graph
    fused_spec: tuple = None           # ("sgd", lr) | ("adam", lr, b1, b2, eps): lets make_pc_step feed the update kernel from
                                       # addresses it already holds (no per-leaf checks / views), see _train._fused_param_update


def _fusable(leaves) -> bool:
    return bool(leaves) and all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() for t in leaves)


def _tables(leaves):
    import ctypes as C
    return (C.c_longlong * len(leaves))(*[t.numel() for t in leaves])


def sgd(learning_rate: float) -> GradientTransformation:
    """optax.sgd(lr): updates = -lr * grads; stateless."""
    def init(params):
        return ()

    def update(updates, state, params=None):
        leaves = tree_leaves(updates)
        new = torch._foreach_mul(leaves, -float(learning_rate)) if leaves else []
        return tree_unflatten_like(updates, list(new)), state

    def fused_flat(p, g, state):
        from . import _lib
        if len(p) != len(g) or not _fusable(p + g):
            return None
        out = _flat.empty_like_list(p)
        _lib.check(_lib.lib.jpcb_optim_sgd(len(p), _tables(p), _lib.ptr_array(p), _lib.ptr_array(g), float(learning_rate),
                                           _lib.ptr_array(out), torch.cuda.current_stream(p[0].device).cuda_stream))
        return out, state

    def fused_apply(params, grads, state):
        res = fused_flat(tree_leaves(params), tree_leaves(grads), state)
        return None if res is None else (tree_unflatten_like(params, res[0]), res[1])
    return GradientTransformation(init, update, fused_apply, fused_flat, float(learning_rate), ("sgd", float(learning_rate)))


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

    def fused_flat(p, g, state):
        from . import _lib
        mu, nu = state["mu"], state["nu"]
        if not (len(p) == len(g) == len(mu) == len(nu)) or not _fusable(p + g + list(mu) + list(nu)):
            return None
        count = state["count"] + 1
        out, mo, vo = (_flat.empty_like_list(p) for _ in range(3))
        _lib.check(_lib.lib.jpcb_optim_adam(len(p), _tables(p), _lib.ptr_array(p), _lib.ptr_array(g), _lib.ptr_array(mu),
                                            _lib.ptr_array(nu), float(learning_rate), float(b1), float(b2), float(eps), count,
                                            _lib.ptr_array(out), _lib.ptr_array(mo), _lib.ptr_array(vo),
                                            torch.cuda.current_stream(p[0].device).cuda_stream))
        return out, {"count": count, "mu": mo, "nu": vo}

    def fused_apply(params, grads, state):
        res = fused_flat(tree_leaves(params), tree_leaves(grads), state)
        return None if res is None else (tree_unflatten_like(params, res[0]), res[1])
    return GradientTransformation(init, update, fused_apply, fused_flat, None,
                                  ("adam", float(learning_rate), float(b1), float(b2), float(eps)))


def apply_updates(model, updates):
    """eqx.apply_updates: leafwise p + u, structure preserved."""
    p, u = tree_leaves(model), tree_leaves(updates)
    new = torch._foreach_add(p, u) if p else []
    return tree_unflatten_like(model, list(new))


def update_and_apply(optim, grads, state, params):
    """`optim.update(grads, state, params)` + `apply_updates(params, updates)` (jpc/_train.py:234-242,
    jpc/_core/_updates.py:189-194); one fused kernel pass when the optimiser offers it and the tensors qualify."""
    fused = getattr(optim, "fused_apply", None)
    if fused is not None:
        res = fused(params, grads, state)
        if res is not None:
            return res
    updates, state = optim.update(grads, state, params)
    return apply_updates(params, updates), state
