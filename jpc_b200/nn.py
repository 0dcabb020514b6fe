"""Minimal model containers with the duck-typed protocol the reference relies on
(`len(model)`, `model[l]`, `model[l].layers[1].weight`, `model[0][1].weight`;
jpc/_core/_energies.py:915-923) plus `make_mlp` / `make_skip_model` / `get_act_fn`
(jpc/_utils.py:13-126).  Weights are torch tensors on the CUDA device; the containers only HOLD
parameters -- all hot-path arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import torch

from ._errors import _check_param_type

_ACT_FNS = ["linear", "tanh", "hard_tanh", "relu", "leaky_relu", "gelu", "selu", "silu"]


class Identity:
    """equinox.nn.Identity stand-in."""
    name = "linear"

    def __call__(self, x, *, key=None):
        return x

    def __repr__(self):
        return "Identity()"


class Activation:
    """A named activation (one of the 8 ids of jpc/_utils.py:13-15).  Calling it evaluates the
    function with torch ops (for user-side evaluation code); the CUDA kernels use the id."""

    def __init__(self, name: str):
        if name not in _ACT_FNS:
            raise ValueError(f"""
                Invalid activation function ID. Options are {_ACT_FNS}.
        """)
        self.name = name

    def __call__(self, x, *, key=None):
        n = self.name
        if n == "linear":
            return x
        if n == "tanh":
            return torch.tanh(x)
        if n == "hard_tanh":
            return torch.clamp(x, -1.0, 1.0)
        if n == "relu":
            return torch.relu(x)
        if n == "leaky_relu":
            return torch.where(x >= 0, x, 0.01 * x)
        if n == "gelu":
            return torch.nn.functional.gelu(x, approximate="tanh")
        if n == "selu":
            return torch.selu(x)
        return x * torch.sigmoid(x)

    def __repr__(self):
        return f"Activation({self.name!r})"


def get_act_fn(name: str) -> Callable:
    """jpc/_utils.py:18-38."""
    if name == "linear":
        return Identity()
    return Activation(name)


class Lambda:
    """equinox.nn.Lambda stand-in: wraps a callable, no parameters."""

    def __init__(self, fn: Callable):
        self.fn = fn

    def __call__(self, x, *, key=None):
        return self.fn(x)

    def __repr__(self):
        return f"Lambda({self.fn!r})"


class Linear:
    """equinox.nn.Linear stand-in: weight [out, in], optional bias [out]; y = W x + b.
    Default init U(-1/sqrt(in), 1/sqrt(in)) for both, as equinox does."""

    def __init__(self, in_features: int, out_features: int, use_bias: bool = True, *, key=None,
                 device=None, dtype=torch.float32):
        gen = _as_generator(key)
        lim = 1.0 / math.sqrt(in_features)
        w = (torch.rand(out_features, in_features, generator=gen, dtype=torch.float64) * 2 - 1) * lim
        self.weight = w.to(dtype).to(device or _default_device())
        self.bias = None
        if use_bias:
            b = (torch.rand(out_features, generator=gen, dtype=torch.float64) * 2 - 1) * lim
            self.bias = b.to(dtype).to(device or _default_device())
        self.in_features, self.out_features, self.use_bias = in_features, out_features, use_bias

    def __call__(self, x, *, key=None):
        y = x @ self.weight.T
        return y if self.bias is None else y + self.bias

    def __repr__(self):
        return f"Linear({self.in_features}, {self.out_features}, use_bias={self.use_bias})"


class Sequential:
    """equinox.nn.Sequential stand-in."""

    def __init__(self, layers: Sequence):
        self.layers = tuple(layers)

    def __getitem__(self, i):
        return self.layers[i]

    def __len__(self):
        return len(self.layers)

    def __iter__(self):
        return iter(self.layers)

    def __call__(self, x, *, key=None):
        for layer in self.layers:
            x = layer(x)
        return x

    def __repr__(self):
        return f"Sequential({list(self.layers)!r})"


def _default_device():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def _as_generator(key) -> Optional[torch.Generator]:
    """`key` plays the role of a jax PRNGKey: an int seed, a torch.Generator, or None."""
    if key is None:
        return None
    if isinstance(key, torch.Generator):
        return key
    return torch.Generator().manual_seed(int(key))


def make_mlp(key, input_dim: int, width: int, depth: int, output_dim: int, act_fn: str,
             use_bias: bool = False, param_type: str = "sp", *, device=None) -> List[Sequential]:
    """jpc/_utils.py:41-108: layer i = Sequential([Lambda(act_i), Linear]), act_0 = Identity,
    i.e. f_l(z) = W_l act(z) + b_l.  `mupc` re-draws the weights from N(0,1)."""
    _check_param_type(param_type)
    gen = _as_generator(key)
    layers = []
    for i in range(depth):
        act_fn_l = Identity() if i == 0 else get_act_fn(act_fn)
        _in = input_dim if i == 0 else width
        _out = output_dim if (i + 1) == depth else width
        linear = Linear(_in, _out, use_bias=use_bias, key=gen, device=device)
        if param_type == "mupc":
            W = torch.randn(_out, _in, generator=gen, dtype=torch.float64)
            linear.weight = W.to(linear.weight.dtype).to(linear.weight.device)
        layers.append(Sequential([Lambda(act_fn_l), linear]))
    return layers


def make_skip_model(depth: int) -> list:
    """jpc/_utils.py:111-126: identity skips at layers 1..depth-2, None elsewhere."""
    skips = [None] * depth
    for l in range(1, depth - 1):
        skips[l] = Lambda(Identity())
    return skips


# ---------------------------------------------------------------------------------------------
# tiny pytree utilities (equinox/jax.tree_util stand-ins for the structures above)
# ---------------------------------------------------------------------------------------------
def _leaves_into(tree, out: list) -> None:
    t = type(tree)  # exact-type dispatch first: these walks run every train step over every layer
    if t is Sequential:
        for c in tree.layers:
            _leaves_into(c, out)
    elif t is Linear:
        out.append(tree.weight)
        if tree.bias is not None:
            out.append(tree.bias)
    elif t is list or t is tuple:
        for c in tree:
            _leaves_into(c, out)
    elif tree is None or t is Lambda:
        return
    elif isinstance(tree, torch.Tensor):
        out.append(tree)
    elif isinstance(tree, (list, tuple)):
        for c in tree:
            _leaves_into(c, out)
    elif isinstance(tree, Sequential):
        for c in tree.layers:
            _leaves_into(c, out)
    elif isinstance(tree, Linear):
        out.append(tree.weight)
        if tree.bias is not None:
            out.append(tree.bias)
    elif isinstance(tree, dict):
        for k in sorted(tree):
            _leaves_into(tree[k], out)


def tree_leaves(tree) -> list:
    """Array leaves (torch tensors) in a deterministic order."""
    out: list = []
    _leaves_into(tree, out)
    return out


def _unflatten(tree, it):
    t = type(tree)
    if t is Sequential:
        new = Sequential.__new__(Sequential)
        new.layers = tuple(_unflatten(c, it) for c in tree.layers)
        return new
    if t is Linear or isinstance(tree, Linear):
        new = Linear.__new__(Linear)
        new.__dict__.update(tree.__dict__)
        new.weight = next(it)
        new.bias = next(it) if tree.bias is not None else None
        return new
    if t is list:
        return [_unflatten(c, it) for c in tree]
    if t is tuple:
        return tuple(_unflatten(c, it) for c in tree)
    if tree is None or t is Lambda:
        return tree
    if isinstance(tree, torch.Tensor):
        return next(it)
    if isinstance(tree, list):
        return [_unflatten(c, it) for c in tree]
    if isinstance(tree, tuple):
        return tuple(_unflatten(c, it) for c in tree)
    if isinstance(tree, Sequential):
        return Sequential([_unflatten(c, it) for c in tree.layers])
    if isinstance(tree, dict):
        return {k: _unflatten(tree[k], it) for k in sorted(tree)}
    return tree  # Identity / Activation / other callables: no leaves


class ModelList(list):
    """A plain list of layers (what the reference's `make_mlp` returns) that can carry the lowering `_Net` made of it
    (`_jpcb`): `make_pc_step` returns its updated model as one, so the next step does not re-derive layer forms, dims and
    parameter addresses from scratch.  Behaves exactly like a list; slicing / copying gives a plain list again."""
    __slots__ = ("_jpcb",)

    def __init__(self, *a):
        super().__init__(*a)
        self._jpcb = None


def _relinear(lin: Linear, weight, bias) -> Linear:
    new = Linear.__new__(Linear)
    d = dict(lin.__dict__)
    d["weight"], d["bias"] = weight, bias
    new.__dict__ = d
    return new


def _unflatten_mlp(model, leaves, i):
    """Fast path of `_unflatten` for a list of Sequential([Lambda, Linear]) layers (jpc.make_mlp models): runs once per
    train step over every layer.  Returns (new model, next leaf index) or None when the structure is anything else."""
    out = ModelList()
    for layer in model:
        if type(layer) is not Sequential:
            return None
        subs = layer.layers
        if len(subs) != 2 or type(subs[0]) is not Lambda or type(subs[1]) is not Linear:
            return None
        lin = subs[1]
        if lin.bias is None:
            new_lin = _relinear(lin, leaves[i], None)
            i += 1
        else:
            new_lin = _relinear(lin, leaves[i], leaves[i + 1])
            i += 2
        seq = Sequential.__new__(Sequential)
        seq.layers = (subs[0], new_lin)
        out.append(seq)
    return out, i


def tree_unflatten_like(tree, leaves):
    """Rebuilds `tree`'s structure around new leaves (taken front to back)."""
    if type(tree) is tuple and len(tree) == 2 and type(tree[0]) in (list, ModelList) and isinstance(leaves, list):
        # (model, skip_model): the make_mlp form with a leaf-free skip model (None / Lambda(Identity) entries)
        sk = tree[1]
        if sk is None or (type(sk) is list and all(e is None or type(e) is Lambda for e in sk)):
            got = _unflatten_mlp(tree[0], leaves, 0)
            if got is not None and got[1] == len(leaves):
                return got[0], (None if sk is None else list(sk))
    elif type(tree) in (list, ModelList) and isinstance(leaves, list):
        got = _unflatten_mlp(tree, leaves, 0)
        if got is not None and got[1] == len(leaves):
            return got[0]
    return _unflatten(tree, iter(leaves))


def tree_map(fn, tree, *rest):
    cols = [tree_leaves(tree)] + [tree_leaves(r) for r in rest]
    new = [fn(*xs) for xs in zip(*cols)]
    return tree_unflatten_like(tree, new)
