"""Host-side mirror of the reference's hot-path function API (jpc/_core/*.py) over the C-ABI.

Every function here keeps the reference's name, argument meaning, return structure and error
behaviour, and does NO arithmetic of its own on the path: it turns the duck-typed model into a
`jpcb_pc_desc`, hands device pointers to libjpc_b200.so, and wraps the outputs back into the
reference's pytrees.  There is no CPU fallback: tensors must live on a CUDA device.

Reference functions mirrored (citations relative to the reference root):
  _get_param_scalings        jpc/_core/_energies.py:879-933   (host arithmetic on python floats)
  pc_energy_fn               jpc/_core/_energies.py:13-158
  init_activities_with_ffwd  jpc/_core/_init.py:12-82
  compute_pc_activity_grad   jpc/_core/_grads.py:95-167
  neg_pc_activity_grad       jpc/_core/_grads.py:22-92
  compute_pc_param_grads     jpc/_core/_grads.py:436-499
  solve_inference            jpc/_core/_infer.py:20-133
  update_pc_activities       jpc/_core/_updates.py:20-110
  update_pc_params           jpc/_core/_updates.py:113-203
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch

from . import _flat, _lib
from ._errors import _check_param_type
from .nn import Identity, Lambda, Linear, ModelList, Sequential, tree_unflatten_like
from .optim import update_and_apply
from .solvers import ConstantStepSize, Euler, Heun, PIDController

# ---------------------------------------------------------------------------------------------
# compute dtype of the GEMM operands (the reference takes it from its arrays' dtype; here the
# master copies are always fp32 and this switch selects fp32 CUDA-core FMA vs bf16 tcgen05)
# ---------------------------------------------------------------------------------------------
_COMPUTE_DTYPE = "f32"


def set_compute_dtype(name: str) -> None:
    """"f32" (fp32 FMA, 1e-5 parity), "bf16" (tcgen05 tensor cores, fp32 accumulate, 2e-2) or "f32x3" (tcgen05 on 3-way
    bf16 split operands, six products per GEMM: fp32 parity at tensor-core speed)."""
    global _COMPUTE_DTYPE
    if name not in ("f32", "bf16", "f32x3"):
        raise ValueError('compute dtype must be "f32", "bf16" or "f32x3"')
    _COMPUTE_DTYPE = name


def get_compute_dtype() -> str:
    return _COMPUTE_DTYPE


# ---------------------------------------------------------------------------------------------
# model -> descriptor
# ---------------------------------------------------------------------------------------------
# torch callables a user may wrap in Lambda instead of jpc_b200.get_act_fn(...).  (torch's gelu defaults to the erf
# form while jax.nn.gelu -- the reference's -- defaults to the tanh approximation, so it is deliberately not mapped.)
_F = torch.nn.functional
_TORCH_ACT_NAMES = {torch.tanh: "tanh", _F.tanh: "tanh", torch.relu: "relu", _F.relu: "relu", _F.hardtanh: "hard_tanh",
                    _F.leaky_relu: "leaky_relu", torch.selu: "selu", _F.selu: "selu", _F.silu: "silu"}


def _act_name(obj):
    """Activation id of a `Lambda(fn)` (or bare callable) in a layer, or None."""
    fn = getattr(obj, "fn", obj)
    name = getattr(fn, "name", None)
    if name is None:
        try:
            name = _TORCH_ACT_NAMES.get(fn)
        except TypeError:
            name = None
    return name if name in _lib.ACT_IDS else None


def _layer_parts(layer) -> Tuple[str, Linear, str]:
    """(input activation id, Linear, output activation id) of one layer.  Two layer forms are on the B200 path:
      Form A  Sequential([Lambda(act), Linear])   f(u) = W act(u) + b    -- what jpc.make_mlp builds (jpc/_utils.py:102-106)
      Form B  Sequential([Linear, Lambda(act)])   f(u) = act(W u + b)    -- the hand-built networks of the reference's
              examples (examples/jpc_from_scratch.ipynb cell 12, examples/discriminative_pc.ipynb cell 8)
    plus a bare Linear (either form with the identity) and Sequential([Lambda, Linear, Lambda]).  Anything else (conv,
    attention, arbitrary callables) has no B200 path, and there is no CPU fallback."""
    if type(layer) is Sequential:  # exact-type fast path: this runs for every layer of every call
        subs = layer.layers
        if len(subs) == 2 and type(subs[1]) is Linear and type(subs[0]) is Lambda:
            name = getattr(subs[0].fn, "name", None)
            if name in _lib.ACT_IDS:
                return name, subs[1], "linear"
    if isinstance(layer, Linear) or (hasattr(layer, "weight") and not hasattr(layer, "layers")):
        return "linear", layer, "linear"
    subs = getattr(layer, "layers", None)
    if subs is not None and 1 <= len(subs) <= 3:
        lin_at = [i for i, m in enumerate(subs) if hasattr(m, "weight")]
        if len(lin_at) == 1:
            i = lin_at[0]
            pre = _act_name(subs[i - 1]) if i >= 1 else "linear"
            post = _act_name(subs[i + 1]) if i + 1 < len(subs) else "linear"
            if pre is not None and post is not None and i <= 1 and len(subs) - i <= 2:
                return pre, subs[i], post
    raise NotImplementedError(
        "jpc_b200 accelerates MLP layers of the forms Sequential([Lambda(act_fn), Linear]) (jpc.make_mlp), "
        "Sequential([Linear, Lambda(act_fn)]) and bare Linear; other layer types have no B200 path and there is no CPU fallback.")


def _skip_flags(skip_model, L: int):
    if skip_model is None:
        return [False] * L
    if len(skip_model) != L:
        raise ValueError("skip_model must have one entry per model layer")
    flags = []
    for s in skip_model:
        if s is None:
            flags.append(False)
            continue
        fn = getattr(s, "fn", s)
        if not isinstance(fn, Identity) and getattr(fn, "name", None) != "linear":
            raise NotImplementedError("only identity skip connections (jpc.make_skip_model) are on the B200 path")
        flags.append(True)
    return flags


def _get_param_scalings(model, input, *, skip_model=None, param_type="sp", gamma=None):
    """jpc/_core/_energies.py:879-933 (python floats; not device work)."""
    L = len(model)
    use_skips = skip_model is not None and any(s is not None for s in skip_model)
    if param_type == "sp":
        scalings = [1.0] * L
    else:
        D = input.shape[1]
        N = _layer_parts(model[0])[1].weight.shape[0]
        a1 = 1 / math.sqrt(D)
        al = 1 / math.sqrt(N * L) if use_skips else 1 / math.sqrt(N)
        aL = 1 / N if param_type == "mupc" else 1 / math.sqrt(N)
        scalings = [a1] + [al] * (L - 2) + [aL]
    if gamma is not None:
        scalings[-1] = scalings[-1] / gamma
    return scalings


class _Shape:
    """Stand-in for an array of which only `.shape` is read."""

    def __init__(self, shape):
        self.shape = tuple(shape)


_DESC_CACHE: dict = {}
_WS_CACHE: dict = {}


def _make_lowering(layers, parts, linears, Ws, bs, dims, dev, w_ptrs, b_ptrs):
    return {"layers": layers, "parts": parts, "linears": linears, "Ws": Ws, "bs": bs,
            "bs_or_none": bs if bs else [None] * len(Ws), "dims": dims, "dev": dev, "w_ptrs": w_ptrs, "b_ptrs": b_ptrs,
            "bare": any(isinstance(l, Linear) for l in layers),
            "acts_t": tuple(p[0] for p in parts), "posts_t": tuple(p[2] for p in parts)}


def relower(net, new_model, new_leaves, addrs):
    """Attaches to `new_model` (a `ModelList` of make_mlp-form layers built by `tree_unflatten_like` from `net`'s model) the
    lowering `make_pc_step` already knows: same layer forms and dims, parameter leaves `new_leaves` (tree order: weight, bias
    per layer) at device addresses `addrs`."""
    if type(new_model) is not ModelList:
        return
    low = net._low
    per = 2 if low["bs"] else 1
    linears = [layer.layers[1] for layer in new_model]
    new_model._jpcb = _make_lowering(list(new_model), [(p[0], lin, p[2]) for p, lin in zip(low["parts"], linears)], linears,
                                     list(new_leaves[0::per]), list(new_leaves[1::per]) if per == 2 else [], low["dims"],
                                     low["dev"], list(addrs[0::per]), list(addrs[1::per]) if per == 2 else [])


class _Net:
    """A model lowered to C-ABI arguments: descriptor + weight/bias pointer arrays."""

    def __init__(self, model, skip_model, x, y, *, param_type, loss_id, gamma, output_energy_scaling,
                 activity_decay, weight_decay, spectral_penalty, solver=0, n_steps=0, dt=0.0, dt_last=0.0,
                 batch_global=None, scaled_free_input=False):
        _check_param_type(param_type)
        if loss_id not in ("mse", "ce"):
            raise ValueError("`mse` and `ce` are the only valid losses.")
        # x is None: unsupervised PC (jpc/_core/_energies.py:86,123-124) -- z_0 = activities[0] is a free activity,
        # every activity list has L + 1 entries and the first layer's prediction carries no scaling
        self.free = x is None
        if self.free and y is None:
            raise ValueError("unsupervised mode (input=None) needs the target to size the batch")
        if self.free and param_type != "sp" and not scaled_free_input:
            raise ValueError("unsupervised mode (input=None) supports param_type='sp' only: the reference's "
                             "`_get_param_scalings` needs the input's width for 'mupc' / 'ntp'")
        L = len(model)
        low = self._lowering(model)
        parts, linears, Ws, bs, dims, dev = low["parts"], low["linears"], low["Ws"], list(low["bs"]), list(low["dims"]), low["dev"]
        # weight_decay / spectral_penalty apply to layers that expose `.weight` directly -- bare Linear layers
        # (jpc/_core/_energies.py:127-143; `> 0` as there); for make_mlp models (Sequential layers) they are no-ops in the
        # reference as well
        self.reg = None
        if low["bare"] and (float(weight_decay) > 0.0 or float(spectral_penalty) > 0.0):
            idx = [i for i, l in enumerate(model) if hasattr(l, "weight") and not hasattr(l, "layers")]
            self.reg = (max(float(weight_decay), 0.0), max(float(spectral_penalty), 0.0), idx)
        self.linears = linears
        self.model, self.skip_model = model, skip_model
        self._low = low
        f32 = torch.float32
        for t in ([x] if x is not None else []) + ([y] if y is not None else []):
            if t.dtype is not f32 or not t.is_contiguous() or t.device != dev:
                raise ValueError("all tensors must be contiguous fp32 on the same CUDA device")
        if bs and len(bs) != L:
            raise NotImplementedError("mixed bias / no-bias layers are not on the B200 path")
        has_bias = [bool(bs)] * L
        if x is not None and x.shape[1] != dims[0]:
            raise ValueError("input feature dim does not match the first layer")
        posts = low["posts_t"]
        self.form_b = any(q != "linear" for q in posts)
        if self.form_b and param_type != "sp":
            # (the reference's _get_param_scalings reads model[0][1].weight, jpc/_core/_energies.py:915: an activation
            #  object for a Sequential([Linear, Lambda]) layer -- it only runs "sp" with that form)
            raise NotImplementedError("layers of the form Sequential([Linear, Lambda(act)]) support param_type='sp' only")
        B = y.shape[0] if self.free else x.shape[0]
        skips = _skip_flags(skip_model, L)
        # (ePC's x=None branch scales the first layer by the width of errors[0] = d_0, jpc/_core/_energies.py:422-427)
        scal_in = x if x is not None else _Shape((B, dims[0]))
        scal = _get_param_scalings(model, scal_in, skip_model=skip_model, param_type=param_type, gamma=gamma)
        self.true_dims = list(dims)
        bs = bs if bs else None
        self.padded = _COMPUTE_DTYPE != "f32" and any(v % 8 for v in dims)  # (both tensor-core paths)
        if self.padded:
            # The tensor-core path needs every feature dim to be a multiple of 8 (TMA row pitch, 16-byte vectors):
            # zero-pad weights, biases, inputs, targets and activities.  Padded entries are zero and stay zero
            # (act(0) = 0 for all 8 activations; zero weight rows / columns), so energies and the un-padded part of
            # every result are unchanged; results are sliced back.  Data layout only -- no arithmetic happens here.
            if loss_id == "ce" and dims[-1] % 8:
                raise NotImplementedError("bf16 path: a cross-entropy output layer must have a multiple of 8 classes "
                                          "(zero-padded logits would enter the softmax)")
            pd = [(v + 7) // 8 * 8 for v in dims]
            P = torch.nn.functional.pad
            Ws = [P(Ws[l], (0, pd[l] - dims[l], 0, pd[l + 1] - dims[l + 1])).contiguous() for l in range(L)]
            if bs is not None:
                bs = [P(bs[l], (0, pd[l + 1] - dims[l + 1])).contiguous() for l in range(L)]
            if x is not None:
                x = P(x, (0, pd[0] - dims[0])).contiguous()
            if y is not None:
                y = P(y, (0, pd[-1] - dims[-1])).contiguous()
            dims = pd
        lam = 1.0 if output_energy_scaling is None else float(output_energy_scaling)
        bg = B if batch_global is None else int(batch_global)
        key = (L, B, bg, tuple(dims), low["acts_t"], posts, all(has_bias), tuple(skips), tuple(scal),
               _COMPUTE_DTYPE, loss_id, lam, float(activity_decay), solver, n_steps, float(dt), float(dt_last), self.free)
        cached = _DESC_CACHE.get(key)
        if cached is None:
            d = _lib.PcDesc()
            d.n_layers, d.batch_local, d.batch_global = L, B, bg
            for i, v in enumerate(dims):
                d.dims[i] = v
            for l in range(L):
                d.act[l] = _lib.ACT_IDS[parts[l][0]]
                d.post_act[l] = _lib.ACT_IDS[posts[l]]
                d.skip[l] = 1 if skips[l] else 0
                d.scalings[l] = scal[l]
            d.has_bias = 1 if all(has_bias) else 0
            d.dtype = _lib.DTYPE_IDS[_COMPUTE_DTYPE]
            d.loss = _lib.LOSS_MSE if loss_id == "mse" else _lib.LOSS_CE
            d.output_energy_scaling = lam
            d.activity_decay = float(activity_decay)
            d.solver, d.n_steps, d.dt, d.dt_last = solver, n_steps, float(dt), float(dt_last)
            d.free_input = 1 if self.free else 0
            ws_bytes = _lib.lib.jpcb_pc_workspace_bytes(d)
            if ws_bytes == 0:
                _lib.check(_probe_error(d))
            cached = (d, ws_bytes)
            if len(_DESC_CACHE) > 256:
                _DESC_CACHE.clear()
            _DESC_CACHE[key] = cached
        self.desc, self.ws_bytes = cached
        self.L, self.B, self.dims, self.dev = L, B, dims, dev
        # widths of the entries of an activity list: [d_1 .. d_L], preceded by d_0 when z_0 is free
        self.adims = dims[0 if self.free else 1:]
        self.true_adims = self.true_dims[0 if self.free else 1:]
        self.has_bias = all(has_bias)
        if self.padded:
            self.w_ptrs = [w.data_ptr() for w in Ws]
            self.b_ptrs = [b_.data_ptr() for b_ in bs] if self.has_bias else None
        else:
            self.w_ptrs, self.b_ptrs = low["w_ptrs"], (low["b_ptrs"] if self.has_bias else None)
        self.W = _lib.ptr_array_from_ints(self.w_ptrs)
        self.b = _lib.ptr_array_from_ints(self.b_ptrs) if self.has_bias else None
        self._keep = (Ws, bs)  # keep tensors alive while pointer arrays exist
        self.x, self.y = x, y   # (zero-padded copies when self.padded)
        self.w_shapes = [tuple(w.shape) for w in Ws]
        self.b_shapes = [tuple(b_.shape) for b_ in bs] if bs is not None else None

    @staticmethod
    def _lowering(model):
        """What `_Net` derives from the model alone -- layer forms, Linear objects, parameter tensors and their device
        addresses, dims -- validated once.  A `ModelList` (what `make_pc_step` returns) carries it to the next call; it is
        trusted only while every layer object and every parameter tensor is still the very object it was made from."""
        low = getattr(model, "_jpcb", None) if type(model) is ModelList else None
        if low is not None and len(low["layers"]) == len(model):
            ok = True
            for layer, ref, lin, w, b_ in zip(model, low["layers"], low["linears"], low["Ws"], low["bs_or_none"]):
                if layer is not ref or lin.weight is not w or lin.bias is not b_:
                    ok = False
                    break
            if ok:
                return low
        parts = [_layer_parts(l) for l in model]
        linears = [p[1] for p in parts]
        dev = linears[0].weight.device
        if dev.type != "cuda":
            raise RuntimeError("jpc_b200 has no CPU path: model weights must be CUDA tensors")
        f32 = torch.float32
        Ws, bs, dims = [], [], [linears[0].weight.shape[1]]
        for l, lin in enumerate(linears):  # one pass per layer
            w, b_ = lin.weight, lin.bias
            if w.dtype is not f32 or not w.is_contiguous() or w.device != dev:
                raise ValueError("all tensors must be contiguous fp32 on the same CUDA device")
            if b_ is not None:
                if b_.dtype is not f32 or not b_.is_contiguous() or b_.device != dev:
                    raise ValueError("all tensors must be contiguous fp32 on the same CUDA device")
                bs.append(b_)
            d_out, d_in = w.shape
            if d_in != dims[l]:
                raise ValueError(f"layer {l}: weight shape {tuple(w.shape)} does not chain from {dims[l]}")
            dims.append(d_out)
            Ws.append(w)
        low = _make_lowering(list(model), parts, linears, Ws, bs, dims, dev, [w.data_ptr() for w in Ws], [b_.data_ptr() for b_ in bs])
        if type(model) is ModelList:
            model._jpcb = low
        return low

    def weight_reg(self, F=None, gW=None):
        """Adds the weight regularisers of the bare Linear layers (jpc/_core/_energies.py:127-143) to the device scalar `F`
        and / or accumulates their gradients into `gW` (per-layer list in the USER's shapes): one library call on the
        layers' own (unpadded) weights.  No-op when the model has none or both coefficients are zero."""
        if self.reg is None:
            return
        import ctypes as C
        wd, sp, idx = self.reg
        Ws = [self.linears[i].weight for i in idx]
        n = len(Ws)
        rows = (C.c_int * n)(*[w.shape[0] for w in Ws])
        cols = (C.c_int * n)(*[w.shape[1] for w in Ws])
        need = _lib.lib.jpcb_pc_weight_reg_workspace_bytes(n, rows, cols)
        ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
        g = None
        if gW is not None:
            for i in idx:
                if not gW[i].is_contiguous():
                    raise ValueError("weight regularisers: gradient buffers must be contiguous")
            g = _lib.ptr_array([gW[i] for i in idx])
        _lib.check(_lib.lib.jpcb_pc_weight_reg(n, _lib.ptr_array(Ws), rows, cols, wd, sp, _lib.ptr(F), 1, g, _lib.ptr(ws), need,
                                               self.stream()))

    def workspace(self) -> torch.Tensor:
        k = (self.dev, self.stream())  # one per stream: calls on different streams may overlap on the device
        ws = _WS_CACHE.get(k)
        if ws is None or ws.numel() < self.ws_bytes:
            ws = torch.empty(max(self.ws_bytes, 1 << 20), dtype=torch.uint8, device=self.dev)
            _WS_CACHE[k] = ws
        return ws

    def stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def with_steps(self, n_steps, dt, dt_last):
        """A copy of the descriptor with another fixed-step schedule (same shapes, same workspace size)."""
        d = _lib.PcDesc.from_buffer_copy(self.desc)
        d.n_steps, d.dt, d.dt_last = int(n_steps), float(dt), float(dt_last)
        return d

    def check_acts(self, activities):
        n = len(self.adims)
        if len(activities) != n:
            raise ValueError(f"expected {n} activity arrays ({'z_0, ' if self.free else ''}one per layer, last = dummy "
                             f"output slot), got {len(activities)}")
        for l, a in enumerate(activities):
            if tuple(a.shape) != (self.B, self.true_adims[l]):
                raise ValueError(f"activities[{l}] has shape {tuple(a.shape)}, expected {(self.B, self.true_adims[l])}")
            if a.device != self.dev or a.dtype != torch.float32 or not a.is_contiguous():
                raise ValueError("activities must be contiguous fp32 CUDA tensors")

    def new_acts(self, lead=()):
        """Output buffers in the layout the library sees (padded feature dims when self.padded): one allocation, carved."""
        lay = _flat.layout([(self.B, w) for w in self.adims])
        return lay.carve(torch.empty(lay.total, dtype=torch.float32, device=self.dev), lead)

    def in_acts(self, activities, clone=False):
        """User activities -> the layout the library sees."""
        if not self.padded:
            return [a.clone() for a in activities] if clone else list(activities)
        P = torch.nn.functional.pad
        return [P(a, (0, self.adims[l] - a.shape[1])).contiguous() for l, a in enumerate(activities)]

    def out_acts(self, bufs):
        """Library-layout activity-shaped results -> the user's shapes."""
        if not self.padded:
            return bufs
        return [b_[:, :self.true_adims[l]].contiguous() for l, b_ in enumerate(bufs)]

    def out_grads(self, gW, gb):
        if not self.padded:
            return gW, gb
        T = self.true_dims
        gW = [g[:T[l + 1], :T[l]].contiguous() for l, g in enumerate(gW)]
        gb = [g[:T[l + 1]].contiguous() for l, g in enumerate(gb)] if gb is not None else None
        return gW, gb


def _probe_error(desc) -> int:
    """workspace_bytes returned 0: re-run validation through a compute entry point to get the code."""
    code = _lib.lib.jpcb_pc_ffwd_init(desc, None, None, None, None, None, None, None, 0, None)
    return code if code != _lib.OK else _lib.EINVAL


def _unpack_params(params):
    if isinstance(params, (tuple, list)) and len(params) == 2 and (params[1] is None or isinstance(params[1], (list, tuple))) \
            and isinstance(params[0], (list, tuple)):
        return params[0], params[1]
    raise ValueError("`params` must be the tuple (model, skip_model)")


def _time_grid(max_t1, dt) -> Tuple[int, float, float]:
    """ConstantStepSize (diffrax): steps of dt from t0=0, the last one clipped to t1."""
    if dt is None or dt <= 0:
        raise ValueError("a fixed-step solve needs dt > 0")
    T = max(1, int(math.ceil(float(max_t1) / float(dt) - 1e-6)))
    last = float(max_t1) - (T - 1) * float(dt)
    return T, float(dt), last


# ---------------------------------------------------------------------------------------------
# reference API
# ---------------------------------------------------------------------------------------------
def init_activities_with_ffwd(model, input, *, skip_model=None, param_type="sp", gamma=None):
    """jpc/_core/_init.py:12-82."""
    net = _Net(model, skip_model, input, None, param_type=param_type, loss_id="mse", gamma=gamma,
               output_energy_scaling=None, activity_decay=0.0, weight_decay=0.0, spectral_penalty=0.0)
    acts = net.new_acts()
    ws = net.workspace()
    _lib.check(_lib.lib.jpcb_pc_ffwd_init(net.desc, net.W, net.b, _lib.ptr(net.x), _lib.ptr_array(acts), None, None,
                                          _lib.ptr(ws), ws.numel(), net.stream()))
    return net.out_acts(acts)


def init_activities_from_normal(key, layer_sizes, mode, batch_size, sigma=0.05, device=None):
    """jpc/_core/_init.py:85-123: z_l ~ N(0, sigma^2) for every layer that is free to vary -- hidden layers and the
    (dummy) output slot in "supervised" mode, also z_0 in "unsupervised" mode.  `key` is a `torch.Generator` or an
    int seed (the reference's is a JAX PRNG key: same law, a different random stream -- a data-side difference,
    not an arithmetic one).  Drawn by torch's device generator directly into device memory."""
    if mode not in ("supervised", "unsupervised"):
        raise ValueError("`mode` must be 'supervised' or 'unsupervised'")
    device = torch.device("cuda" if device is None else device)
    if isinstance(key, torch.Generator):
        gen = key
    else:
        gen = torch.Generator(device=device)
        gen.manual_seed(int(key))
    start_l = 0 if mode == "unsupervised" else 1
    return [float(sigma) * torch.randn(int(batch_size), int(layer_sizes[l]), generator=gen, device=gen.device,
                                       dtype=torch.float32).to(device)
            for l in range(start_l, len(layer_sizes))]


def pc_energy_fn(params, activities, y, *, x=None, loss="mse", param_type="sp", weight_decay=0.0,
                 spectral_penalty=0.0, activity_decay=0.0, record_layers=False, gamma=None,
                 output_energy_scaling=None):
    """jpc/_core/_energies.py:13-158.  Scalar total energy, or the per-layer array [output, hidden
    1..L-2, first] / B when record_layers=True."""
    model, skip_model = _unpack_params(params)
    net = _Net(model, skip_model, x, y, param_type=param_type, loss_id=loss, gamma=gamma,
               output_energy_scaling=output_energy_scaling, activity_decay=activity_decay,
               weight_decay=weight_decay, spectral_penalty=spectral_penalty)
    net.check_acts(activities)
    ws = net.workspace()
    if record_layers:
        E = torch.empty(net.L, dtype=torch.float32, device=net.dev)
        _lib.check(_lib.lib.jpcb_pc_energy(net.desc, net.W, net.b, _lib.ptr_array(net.in_acts(activities)), _lib.ptr(net.x), _lib.ptr(net.y),
                                           _lib.ptr(E), _lib.ptr(ws), ws.numel(), net.stream()))
        return E
    F = torch.empty((), dtype=torch.float32, device=net.dev)
    g = net.new_acts()
    _lib.check(_lib.lib.jpcb_pc_activity_grad(net.desc, net.W, net.b, _lib.ptr_array(net.in_acts(activities)), _lib.ptr(net.x),
                                              _lib.ptr(net.y), _lib.ptr(F), _lib.ptr_array(g), _lib.ptr(ws), ws.numel(), net.stream()))
    net.weight_reg(F=F)
    return F


def compute_pc_activity_grad(params, activities, y, *, x, loss_id="mse", param_type="sp", weight_decay=0.0,
                             spectral_penalty=0.0, activity_decay=0.0, gamma=None, output_energy_scaling=None,
                             batch_global=None):
    """jpc/_core/_grads.py:95-167: (energy, dF/dactivities).  `batch_global` (not in the reference signature): the 1/B
    of the energy when `activities` is one rank's shard of a data-parallel batch."""
    model, skip_model = _unpack_params(params)
    net = _Net(model, skip_model, x, y, param_type=param_type, loss_id=loss_id, gamma=gamma,
               output_energy_scaling=output_energy_scaling, activity_decay=activity_decay,
               weight_decay=weight_decay, spectral_penalty=spectral_penalty, batch_global=batch_global)
    net.check_acts(activities)
    ws = net.workspace()
    F = torch.empty((), dtype=torch.float32, device=net.dev)
    g = net.new_acts()
    _lib.check(_lib.lib.jpcb_pc_activity_grad(net.desc, net.W, net.b, _lib.ptr_array(net.in_acts(activities)), _lib.ptr(net.x),
                                              _lib.ptr(net.y), _lib.ptr(F), _lib.ptr_array(g), _lib.ptr(ws), ws.numel(), net.stream()))
    net.weight_reg(F=F)   # (the energy value includes the weight regularisers; the activity gradients do not depend on them)
    return F, net.out_acts(g)


def neg_pc_activity_grad(t, activities, args):
    """jpc/_core/_grads.py:22-92: the ODE vector field -dF/dz with the reference's 11-tuple args."""
    params, y, x, loss_id, param_type, weight_decay, spectral_penalty, activity_decay, gamma, output_energy_scaling, _ = args
    _, g = compute_pc_activity_grad(params, activities, y, x=x, loss_id=loss_id, param_type=param_type,
                                    weight_decay=weight_decay, spectral_penalty=spectral_penalty,
                                    activity_decay=activity_decay, gamma=gamma,
                                    output_energy_scaling=output_energy_scaling)
    torch._foreach_neg_(g)
    return g


def _param_grads_raw(net: _Net, activities, want_energies=False, arena=None):
    """Runs jpcb_pc_param_grads; returns (gW list, gb list or None, E_layers or None)."""
    f32 = dict(dtype=torch.float32, device=net.dev)
    gW = [torch.empty(sh, **f32) for sh in net.w_shapes]
    gb = [torch.empty(sh, **f32) for sh in net.b_shapes] if net.has_bias else None
    E = torch.empty(net.L, dtype=torch.float32, device=net.dev) if want_energies else None
    ws = net.workspace()
    _lib.check(_lib.lib.jpcb_pc_param_grads(net.desc, net.W, net.b, _lib.ptr_array(net.in_acts(activities)), _lib.ptr(net.x),
                                            _lib.ptr(net.y), _lib.ptr_array(gW), _lib.ptr_array(gb) if gb else None, _lib.ptr(E),
                                            _lib.ptr(ws), ws.numel(), net.stream()))
    gW, gb = net.out_grads(gW, gb)
    net.weight_reg(gW=gW)
    return gW, gb, E


def _grads_pytree(net: _Net, gW, gb):
    """(model_grads, skip_grads) with the model's structure (jpc/_core/_grads.py:487-499): the skip
    model (Lambda(Identity) entries) has no array leaves, so its gradient pytree is all-None."""
    leaves = []
    for l in range(net.L):
        leaves.append(gW[l])
        if gb is not None:
            leaves.append(gb[l])
    model_grads = tree_unflatten_like(net.model, leaves)
    skip_grads = None if net.skip_model is None else [None] * len(net.skip_model)
    return model_grads, skip_grads


def compute_pc_param_grads(params, activities, y, *, x=None, loss_id="mse", param_type="sp", weight_decay=0.0,
                           spectral_penalty=0.0, activity_decay=0.0, gamma=None, output_energy_scaling=None):
    """jpc/_core/_grads.py:436-499."""
    model, skip_model = _unpack_params(params)
    net = _Net(model, skip_model, x, y, param_type=param_type, loss_id=loss_id, gamma=gamma,
               output_energy_scaling=output_energy_scaling, activity_decay=activity_decay,
               weight_decay=weight_decay, spectral_penalty=spectral_penalty)
    net.check_acts(activities)
    gW, gb, _ = _param_grads_raw(net, activities)
    return _grads_pytree(net, gW, gb)


def _solver_id(solver) -> int:
    if isinstance(solver, Euler):
        return _lib.SOLVER_EULER
    if isinstance(solver, Heun):
        return _lib.SOLVER_HEUN
    raise NotImplementedError(f"solver {solver!r} is not on the B200 path (Euler and Heun are)")


MAX_STEPS = 4096  # diffrax.diffeqsolve default


def _rms(tensors, reduce_over_ranks=False) -> float:
    """optimistix.rms_norm over a list of arrays (host read-back of one scalar; controller logic only).  With
    `reduce_over_ranks` the norm is that of the GLOBAL data-parallel batch: sums of squares and element counts are
    all-reduced, so every rank gets the same number (the controller's decisions must not diverge between ranks)."""
    tot = sum(torch.sum(t.double() ** 2) for t in tensors)
    n = sum(t.numel() for t in tensors)
    if reduce_over_ranks:
        import torch.distributed as dist
        buf = torch.stack([tot, torch.tensor(float(n), dtype=torch.float64, device=tot.device)])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        tot, n = buf[0], float(buf[1])
    return math.sqrt(float(tot) / n)


class _SaveAt:
    """The `SaveAt` of jpc/_core/_infer.py:103-107, restated (diffrax is third-party, recalled: SURVEY App. B).
    * neither flag: `SaveAt(t1=True)` -> one row, the final state: leaves (1, B, d).
    * `record_iters`: `SaveAt(t1=True, steps=True)` -> one row per accepted step, leaves (4096, B, d) padded with
      inf after the last step (diffrax `max_steps`; the final state is the last finite row, `get_t_max`).
    * `record_every`: `SaveAt(t1=True, ts=arange(0, max_t1, record_every))` -> one row per requested time, taken
      by linear interpolation inside the accepted step that contains it (diffrax LocalLinearInterpolation, the
      dense output of both Euler and Heun), inf where the solve stopped early, then one row with the final state.
    Rows are device copies of the activity buffers made between C-ABI calls: bookkeeping, no arithmetic other
    than the interpolation `torch.lerp`."""

    def __init__(self, net, max_t1, record_iters, record_every):
        self.net = net
        self.ts = None
        if record_every is not None:
            self.ts = torch.arange(0, max_t1, record_every).tolist()
            rows = len(self.ts) + 1
        elif record_iters:
            rows = MAX_STEPS
        else:
            rows = 1
        self.steps = bool(record_iters) and self.ts is None
        self.ys = [torch.full((rows, net.B, w), float("inf"), dtype=torch.float32, device=net.dev) for w in net.true_adims]
        self.i = 0

    @property
    def needs_prev(self):
        return self.ts is not None

    def _put(self, row, bufs, prev=None, w=None):
        for l, y in enumerate(self.ys):
            d = self.net.true_adims[l]
            if prev is None:
                y[row].copy_(bufs[l][:, :d])
            else:
                torch.lerp(prev[l][:, :d], bufs[l][:, :d], w, out=y[row])

    def step(self, t0, prev, t1, cur):
        """One accepted step from (t0, prev) to (t1, cur); buffers in the library's (possibly padded) layout."""
        if self.ts is not None:
            while self.i < len(self.ts) and self.ts[self.i] <= t1:
                self._put(self.i, cur, prev, (self.ts[self.i] - t0) / (t1 - t0))
                self.i += 1
        elif self.steps:
            self._put(self.i, cur)
            self.i += 1

    def finish(self, cur):
        if self.ts is not None:
            # diffrax writes the t1 row at the current save index: right after the last `ts` row it filled (the rows
            # beyond stay inf when a steady-state event ended the solve early; the last row when it ran to t1)
            self._put(self.i, cur)
        elif not self.steps:
            self._put(-1, cur)
        return self.ys


def _solve_adaptive(params, activities, output, input, loss_id, param_type, solver, max_t1, dt0, ctrl, weight_decay,
                    spectral_penalty, activity_decay, gamma, output_energy_scaling, batch_global=None, record_iters=False,
                    record_every=None):
    """solve_inference with diffrax.Heun + PIDController (the reference's default, jpc/_core/_infer.py:28-33):
    every trial step (two gradient evaluations, candidate state, embedded-error and activity sums of squares) is
    one C-ABI call; accept/reject and the next step size are the controller's scalar arithmetic on the host
    (solvers.PIDController), one 8-byte read-back per trial step.  The steady-state Event of the reference
    (`_infer.py:136-145`) is evaluated after every accepted step."""
    if not isinstance(solver, Heun):
        if isinstance(solver, Euler):
            raise RuntimeError("Cannot use adaptive step sizes with a solver that does not provide error estimates.")
        raise NotImplementedError(f"solver {solver!r} is not on the B200 path (Euler and Heun are)")
    model, skip_model = _unpack_params(params)
    net = _Net(model, skip_model, input, output, param_type=param_type, loss_id=loss_id, gamma=gamma,
               output_energy_scaling=output_energy_scaling, activity_decay=activity_decay, weight_decay=weight_decay,
               spectral_penalty=spectral_penalty, solver=_lib.SOLVER_HEUN, n_steps=1, dt=0.0, dt_last=0.0,
               batch_global=batch_global)
    net.check_acts(activities)
    order = 2.0  # diffrax: error order of Heun (2nd order with an embedded 1st-order estimate)
    kw_out = output  # (user-shaped tensors for the two gradient evaluations of the initial-step heuristic)
    kw = dict(x=input, loss_id=loss_id, param_type=param_type, weight_decay=weight_decay, spectral_penalty=spectral_penalty,
              activity_decay=activity_decay, gamma=gamma, output_energy_scaling=output_energy_scaling)
    t1 = float(max_t1)
    A = net.in_acts(activities, clone=True)
    B_ = [torch.empty_like(a) for a in A]
    ws = net.workspace()
    out2 = torch.zeros(2, dtype=torch.float32, device=net.dev)
    input, output = net.x, net.y
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    dp = world > 1
    # the controller's norms are over the GLOBAL batch: global element count (shards may be uneven; padding
    # entries are zero and add nothing to the sums), global sums of squares, the gradient's 1/B that of the global batch --
    # so every rank computes the same step sizes and takes the same accept / reject decisions
    numel = sum(a.numel() for a in activities)
    if dp:
        nbuf = torch.tensor([float(numel)], dtype=torch.float64, device=net.dev)
        dist.all_reduce(nbuf, op=dist.ReduceOp.SUM)
        numel = int(nbuf.item())
    if dt0 is None:
        # Hairer / Wanner initial step, as diffrax's PIDController.init does when dt0 is None
        A0 = [a.contiguous() for a in activities]
        _, g0 = compute_pc_activity_grad(params, A0, kw_out, batch_global=net.desc.batch_global, **kw)
        f0 = [-g for g in g0]
        scale = [ctrl.atol + a.abs() * ctrl.rtol for a in A0]
        d0 = _rms([a / s for a, s in zip(A0, scale)], dp)
        d1 = _rms([f / s for f, s in zip(f0, scale)], dp)
        small = d0 < 1e-5 or d1 < 1e-5
        h0 = 1e-6 if small else 0.01 * d0 / d1
        y1 = [a + h0 * f for a, f in zip(A0, f0)]
        _, g1 = compute_pc_activity_grad(params, y1, kw_out, batch_global=net.desc.batch_global, **kw)
        d2 = _rms([(-gb - f) / s for gb, f, s in zip(g1, f0, scale)], dp) / h0
        max_d = max(d1, d2)
        h1 = max(1e-6, h0 * 1e-3) if max_d <= 1e-15 else (0.01 / max_d) ** (1.0 / order)
        dt = min(100.0 * h0, h1)
    else:
        dt = float(dt0)
    t, steps, state = 0.0, 0, ctrl.init_state()
    stats = {"accepted": 0, "rejected": 0, "event": False}
    save = _SaveAt(net, max_t1, record_iters, record_every)
    while t < t1:
        if steps >= MAX_STEPS:
            raise RuntimeError("The maximum number of solver steps was reached. Try increasing `max_steps`.")
        h = min(dt, t1 - t)
        _lib.check(_lib.lib.jpcb_pc_heun_trial(net.desc, net.W, net.b, _lib.ptr_array(A), _lib.ptr(input), _lib.ptr(output),
                                               h, ctrl.rtol, ctrl.atol, _lib.ptr_array(B_), _lib.ptr(out2), _lib.ptr(ws),
                                               ws.numel(), net.stream()))
        if world > 1:  # the controller's norms are global over the batch: one scalar all-reduce per trial step
            dist.all_reduce(out2, op=dist.ReduceOp.SUM)
        err_ss, act_ss = out2.tolist()  # the one host read-back per trial step
        keep, dt, state = ctrl.adapt(h, math.sqrt(err_ss / numel), order, state)
        steps += 1
        if keep:
            t_prev, t = t, (t1 if t1 - (t + h) < 1e-6 * max(1.0, t1) else t + h)
            A, B_ = B_, A
            save.step(t_prev, B_, t, A)
            stats["accepted"] += 1
            rms = math.sqrt(act_ss / numel)
            if rms < ctrl.atol + ctrl.rtol * rms:  # steady_state_event_with_timeout
                stats["event"] = True
                break
        else:
            stats["rejected"] += 1
    solve_inference.last_stats = stats
    return save.finish(A)


def solve_inference(params, activities, output, *, input=None, loss_id="mse", param_type="sp", solver=None,
                    max_t1=20, dt=None, stepsize_controller=None, weight_decay=0.0, spectral_penalty=0.0,
                    activity_decay=0.0, gamma=None, output_energy_scaling=None, record_iters=False,
                    record_every=None, batch_global=None):
    """jpc/_core/_infer.py:20-133.  Fixed-step solvers (ConstantStepSize): all T steps x L layers run on the
    device without returning to the host.  DEVIATION: the reference's steady-state Event (`_infer.py:136-145`) also
    applies to fixed steps; with ConstantStepSize it falls back to atol = rtol = 1e-3 and stops the solve after a step
    once rms(ACTIVITIES) < 1e-3 (1 + rms) -- i.e. only for activities that are all ~1e-3 or smaller (zero or tiny
    initial activities), never for feed-forward-initialised networks.  The fused loop does not evaluate it: such a solve
    runs all T steps here.  PIDController (the default): controller-driven trial steps, see `_solve_adaptive`.  Returns leaves
    with the reference's leading time axis: (1, B, d) by default (SaveAt(t1=True)), (4096, B, d) inf-padded with
    `record_iters`, (len(ts) + 1, B, d) with `record_every` (see `_SaveAt`).  `batch_global` (not in the reference
    signature) is the 1/B of the energy when `activities` is one rank's shard of a data-parallel batch."""
    solver = Heun() if solver is None else solver
    stepsize_controller = PIDController(rtol=1e-3, atol=1e-3) if stepsize_controller is None else stepsize_controller
    if isinstance(stepsize_controller, PIDController):
        return _solve_adaptive(params, activities, output, input, loss_id, param_type, solver, max_t1, dt, stepsize_controller,
                               weight_decay, spectral_penalty, activity_decay, gamma, output_energy_scaling, batch_global,
                               record_iters, record_every)
    if not isinstance(stepsize_controller, ConstantStepSize):
        raise NotImplementedError(f"unsupported step size controller {stepsize_controller!r}")
    model, skip_model = _unpack_params(params)
    T, h, h_last = _time_grid(max_t1, dt)
    if T > 4096:
        raise RuntimeError("The maximum number of solver steps (4096) was reached.")  # diffrax max_steps
    net = _Net(model, skip_model, input, output, param_type=param_type, loss_id=loss_id, gamma=gamma,
               output_energy_scaling=output_energy_scaling, activity_decay=activity_decay,
               weight_decay=weight_decay, spectral_penalty=spectral_penalty,
               solver=_solver_id(solver), n_steps=T, dt=h, dt_last=h_last, batch_global=batch_global)
    net.check_acts(activities)
    A = net.in_acts(activities, clone=True)  # functional API: inputs are never mutated
    ws = net.workspace()
    if not record_iters and record_every is None:
        _lib.check(_lib.lib.jpcb_pc_infer(net.desc, net.W, net.b, _lib.ptr_array(A), _lib.ptr(net.x), _lib.ptr(net.y),
                                          None, _lib.ptr(ws), ws.numel(), net.stream()))
        return [a.unsqueeze(0) for a in net.out_acts(A)]
    # recording: the same entry point, one solver step per call, rows copied out in between (`_SaveAt`)
    save = _SaveAt(net, max_t1, record_iters, record_every)
    one, last = net.with_steps(1, h, h), net.with_steps(1, h_last, h_last)
    prev = [torch.empty_like(a) for a in A] if save.needs_prev else None
    for i in range(T):
        if prev is not None:
            for p_, a in zip(prev, A):
                p_.copy_(a)
        d = last if i == T - 1 else one
        _lib.check(_lib.lib.jpcb_pc_infer(d, net.W, net.b, _lib.ptr_array(A), _lib.ptr(net.x), _lib.ptr(net.y),
                                          None, _lib.ptr(ws), ws.numel(), net.stream()))
        save.step(i * h, prev, float(max_t1) if i == T - 1 else (i + 1) * h, A)
    return save.finish(A)


def update_pc_activities(params, activities, optim, opt_state, output, *, input=None, loss_id="mse", param_type="sp",
                         weight_decay=0.0, spectral_penalty=0.0, activity_decay=0.0, gamma=None,
                         output_energy_scaling=None):
    """jpc/_core/_updates.py:20-110."""
    energy, grads = compute_pc_activity_grad(params, activities, output, x=input, loss_id=loss_id, param_type=param_type,
                                             weight_decay=weight_decay, spectral_penalty=spectral_penalty,
                                             activity_decay=activity_decay, gamma=gamma,
                                             output_energy_scaling=output_energy_scaling)
    activities, opt_state = update_and_apply(optim, grads, opt_state, activities)
    return {"energy": energy, "activities": activities, "grads": grads, "opt_state": opt_state}


def update_pc_params(params, activities, optim, opt_state, output, *, input=None, loss_id="mse", param_type="sp",
                     weight_decay=0.0, spectral_penalty=0.0, activity_decay=0.0, gamma=None, output_energy_scaling=None):
    """jpc/_core/_updates.py:113-203."""
    grads = compute_pc_param_grads(params, activities, output, x=input, loss_id=loss_id, param_type=param_type,
                                   weight_decay=weight_decay, spectral_penalty=spectral_penalty,
                                   activity_decay=activity_decay, gamma=gamma,
                                   output_energy_scaling=output_energy_scaling)
    (model, skip_model), opt_state = update_and_apply(optim, grads, opt_state, tuple(params))
    return {"model": model, "skip_model": skip_model, "grads": grads, "opt_state": opt_state}
