"""Host-side mirror of the reference's bidirectional-PC functions over the jpcb_bpc_* C-ABI.

  bpc_energy_fn              jpc/_core/_energies.py:246-357
  compute_bpc_activity_grad  jpc/_core/_grads.py:170-226
  compute_bpc_param_grads    jpc/_core/_grads.py:544-612
  update_bpc_activities      jpc/_core/_updates.py:206-281
  update_bpc_params          jpc/_core/_updates.py:284-372
  solve_bpc_inference        (no reference twin: the fused equivalent of the reference's Python loop
                              `for _ in range(T): update_bpc_activities(..., optax.sgd(dt))`,
                              examples/bidirectional_pc.ipynb cell 8 -- all T steps stay on the device)

Same names, arguments, return structures and exceptions as the reference; no arithmetic here.
"""
from __future__ import annotations

import torch

from . import _lib
from . import _core, _flat
from ._core import _layer_parts, _skip_flags, _WS_CACHE
from ._errors import _check_param_type
from .nn import tree_unflatten_like
from .optim import update_and_apply


_BPC_DESC_CACHE: dict = {}


class _BpcNet:
    """(top_down, bottom_up) lowered to a jpcb_bpc_desc + pointer arrays."""

    def __init__(self, top_down_model, bottom_up_model, x, y, *, skip_model, param_type, backward_energy_weight,
                 forward_energy_weight, n_steps=0, dt=0.0, batch_global=None):
        _check_param_type(param_type)  # validated; bpc_energy_fn applies no scalings (_energies.py:314)
        # x = None (jpc/_core/_energies.py:324): activities[0] stands in for the input -- desc.td.free_input, `x` pointer NULL.
        self.free = x is None
        if len(top_down_model) != len(bottom_up_model):
            raise ValueError("top_down_model and bottom_up_model must have the same number of layers")
        n = len(top_down_model)
        H = n - 1
        td = [_layer_parts(l) for l in top_down_model]
        bu = [_layer_parts(l) for l in bottom_up_model]
        if any(q[2] != "linear" for q in td + bu):
            raise NotImplementedError("bPC: layers of the form Sequential([Linear, Lambda(act)]) are not on the B200 path")
        self.td_lin, self.bu_lin = [p[1] for p in td], [p[1] for p in bu]
        self.top_down_model, self.bottom_up_model = top_down_model, bottom_up_model
        Ws, Vs = [l.weight for l in self.td_lin], [l.weight for l in self.bu_lin]
        dev = Ws[0].device
        if dev.type != "cuda":
            raise RuntimeError("jpc_b200 has no CPU path: model weights must be CUDA tensors")
        for t in Ws + Vs + ([y] if self.free else [x, y]):
            if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("all tensors must be contiguous fp32 on the same CUDA device")
        dims = [Ws[0].shape[1]] + [w.shape[0] for w in Ws]
        self.true_dims = list(dims)
        for j in range(n):
            if Ws[j].shape[1] != dims[j]:
                raise ValueError(f"top-down layer {j}: weight shape {tuple(Ws[j].shape)} does not chain from {dims[j]}")
            if tuple(Vs[j].shape) != (dims[j], dims[j + 1]):
                raise ValueError(f"bottom-up layer {j}: weight shape {tuple(Vs[j].shape)}, expected {(dims[j], dims[j + 1])}")
        if self.free:
            if dims[0] != dims[1]:
                raise ValueError(f"bPC with x=None feeds activities[0] (width {dims[1]}) to top_down_model[0] (expects {dims[0]})")
            if y.shape[1] != dims[-1]:
                raise ValueError("output shape does not match the models")
        elif x.shape[1] != dims[0] or y.shape[1] != dims[-1] or x.shape[0] != y.shape[0]:
            raise ValueError("input / output shapes do not match the models")
        td_bias = [l.bias is not None for l in self.td_lin]
        bu_bias = [l.bias is not None for l in self.bu_lin]
        if (any(td_bias) and not all(td_bias)) or (any(bu_bias) and not all(bu_bias)):
            raise NotImplementedError("mixed bias / no-bias layers are not on the B200 path")
        skips = _skip_flags(skip_model, n)
        B = y.shape[0]
        self.dtype = _core.get_compute_dtype()
        self.bf16 = self.dtype != "f32"  # either tensor-core path: same layout rules (dims / batch multiples of 8)
        bWs = [l.bias for l in self.td_lin] if all(td_bias) else None
        bVs = [l.bias for l in self.bu_lin] if all(bu_bias) else None
        if self.bf16:
            # The tensor-core path needs every feature dim to be a multiple of 8 (TMA row pitch, 16-byte vectors):
            # zero-pad.  Padded weights / inputs / activities are zero and stay zero (act(0) = 0 for all 8
            # activations; zero weight rows and columns), so energies and the un-padded parts of every result are
            # unchanged; results are sliced back.  (Data layout only -- no arithmetic happens here.)
            pd = [(v + 7) // 8 * 8 for v in dims]

            def P(t, pads):  # (only the tensors that actually need padding are copied: this runs every call)
                return torch.nn.functional.pad(t, pads).contiguous() if any(pads) else t
            Ws = [P(Ws[j], (0, pd[j] - dims[j], 0, pd[j + 1] - dims[j + 1])) for j in range(n)]
            Vs = [P(Vs[j], (0, pd[j + 1] - dims[j + 1], 0, pd[j] - dims[j])) for j in range(n)]
            if bWs is not None:
                bWs = [P(bWs[j], (0, pd[j + 1] - dims[j + 1])) for j in range(n)]
            if bVs is not None:
                bVs = [P(bVs[j], (0, pd[j] - dims[j])) for j in range(n)]
            if not self.free:
                x = P(x, (0, pd[0] - dims[0]))
            y = P(y, (0, pd[-1] - dims[-1]))
            dims = pd
        bg = B if batch_global is None else int(batch_global)
        key = (n, B, bg, tuple(dims), tuple(p[0] for p in td), tuple(p[0] for p in bu), tuple(skips), all(td_bias), all(bu_bias),
               self.dtype, int(n_steps), float(dt), float(forward_energy_weight), float(backward_energy_weight), self.free)
        cached = _BPC_DESC_CACHE.get(key)
        if cached is None:
            d = _lib.BpcDesc()
            d.td.n_layers, d.td.batch_local, d.td.batch_global = n, B, bg
            for i, v in enumerate(dims):
                d.td.dims[i] = v
            for j in range(n):
                d.td.act[j] = _lib.ACT_IDS[td[j][0]]
                d.bu_act[j] = _lib.ACT_IDS[bu[j][0]]
                d.td.skip[j] = 1 if skips[j] else 0
                d.td.scalings[j] = 1.0
            d.td.has_bias, d.bu_has_bias = int(all(td_bias)), int(all(bu_bias))
            d.td.dtype, d.td.loss = _lib.DTYPE_IDS[self.dtype], _lib.LOSS_MSE
            d.td.output_energy_scaling, d.td.activity_decay = 1.0, 0.0
            d.td.free_input = int(self.free)
            d.td.solver, d.td.n_steps, d.td.dt, d.td.dt_last = _lib.SOLVER_EULER, int(n_steps), float(dt), float(dt)
            d.forward_energy_weight, d.backward_energy_weight = float(forward_energy_weight), float(backward_energy_weight)
            ws_bytes = _lib.lib.jpcb_bpc_workspace_bytes(d)
            if ws_bytes == 0:
                _lib.check(_lib.lib.jpcb_bpc_energy(d, None, None, None, None, None, None, None, None, None, 0, None) or _lib.EINVAL)
            if len(_BPC_DESC_CACHE) > 128:
                _BPC_DESC_CACHE.clear()
            cached = _BPC_DESC_CACHE[key] = (d, ws_bytes)
        self.desc, self.ws_bytes = cached
        self.H, self.B, self.dims, self.dev = H, B, dims, dev
        self.td_bias, self.bu_bias = all(td_bias), all(bu_bias)
        self.W, self.V = _lib.ptr_array(Ws), _lib.ptr_array(Vs)
        self.bW = _lib.ptr_array(bWs) if self.td_bias else None
        self.bV = _lib.ptr_array(bVs) if self.bu_bias else None
        self._keep = (Ws, Vs, bWs, bVs)
        self.x, self.y = x, y

    def workspace(self):
        k = (self.dev, self.stream())  # one per stream (shared with the PC entry points on that stream)
        ws = _WS_CACHE.get(k)
        if ws is None or ws.numel() < self.ws_bytes:
            ws = torch.empty(max(self.ws_bytes, 1 << 20), dtype=torch.uint8, device=self.dev)
            _WS_CACHE[k] = ws
        return ws

    def stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def check_acts(self, activities):
        """H+1 arrays: z_0..z_{H-1} and the unused last slot (activities[-2] must be z_{H-1},
        jpc/_core/_energies.py:331)."""
        if len(activities) != self.H + 1:
            raise ValueError(f"expected {self.H + 1} activity arrays (z_0..z_{self.H - 1} + the unused output slot), got {len(activities)}")
        for k, a in enumerate(activities[:-1]):
            if tuple(a.shape) != (self.B, self.true_dims[k + 1]):
                raise ValueError(f"activities[{k}] has shape {tuple(a.shape)}, expected {(self.B, self.true_dims[k + 1])}")
        for a in activities:
            if a.device != self.dev or a.dtype != torch.float32 or not a.is_contiguous():
                raise ValueError("activities must be contiguous fp32 CUDA tensors")

    def pad_acts(self, activities, clone=False):
        """Activities in the layout the library sees (zero-padded feature dims on the bf16 path)."""
        if not self.bf16:
            return [a.clone() for a in activities] if clone else list(activities)
        P = torch.nn.functional.pad
        out = [P(a, (0, self.dims[k + 1] - a.shape[1])).contiguous() if self.dims[k + 1] != a.shape[1] else (a.clone() if clone else a)
               for k, a in enumerate(activities[:-1])]
        out.append(torch.zeros(self.B, self.dims[-1], dtype=torch.float32, device=self.dev))
        return out

    def unpad_acts(self, padded, like):
        if not self.bf16:
            return padded
        return [p[:, :a.shape[1]].contiguous() if p.shape[1] >= a.shape[1] else a.clone() for p, a in zip(padded, like)]

    def args(self):
        return (self.desc, self.W, self.bW, self.V, self.bV)


def _net(top_down_model, bottom_up_model, x, y, skip_model, param_type, backward_energy_weight, forward_energy_weight, **kw):
    return _BpcNet(top_down_model, bottom_up_model, x, y, skip_model=skip_model, param_type=param_type,
                   backward_energy_weight=backward_energy_weight, forward_energy_weight=forward_energy_weight, **kw)


def bpc_energy_fn(top_down_model, bottom_up_model, activities, y, *, x=None, skip_model=None, param_type="sp",
                  record_layers=False, backward_energy_weight=1.0, forward_energy_weight=1.0):
    """jpc/_core/_energies.py:246-357."""
    net = _net(top_down_model, bottom_up_model, x, y, skip_model, param_type, backward_energy_weight, forward_energy_weight)
    net.check_acts(activities)
    ws = net.workspace()
    E = torch.empty(2 * (net.H + 1), dtype=torch.float32, device=net.dev)
    _lib.check(_lib.lib.jpcb_bpc_energy(*net.args(), _lib.ptr_array(net.pad_acts(activities)), _lib.ptr(net.x), _lib.ptr(net.y), _lib.ptr(E),
                                        _lib.ptr(ws), ws.numel(), net.stream()))
    return E if record_layers else E.sum()


def compute_bpc_activity_grad(top_down_model, bottom_up_model, activities, y, *, x=None, skip_model=None, param_type="sp",
                              backward_energy_weight=1.0, forward_energy_weight=1.0):
    """jpc/_core/_grads.py:170-226: (energy, dF/dactivities)."""
    net = _net(top_down_model, bottom_up_model, x, y, skip_model, param_type, backward_energy_weight, forward_energy_weight)
    net.check_acts(activities)
    ws = net.workspace()
    F = torch.empty((), dtype=torch.float32, device=net.dev)
    pa = net.pad_acts(activities)
    # (the unused last slot may have any shape in the reference; its gradient is zero)
    g = [torch.empty_like(a) for a in pa[:-1]] + [torch.empty(net.B, net.dims[-1], dtype=torch.float32, device=net.dev)]
    _lib.check(_lib.lib.jpcb_bpc_activity_grad(*net.args(), _lib.ptr_array(pa), _lib.ptr(net.x), _lib.ptr(net.y), _lib.ptr(F),
                                               _lib.ptr_array(g), _lib.ptr(ws), ws.numel(), net.stream()))
    return F, net.unpad_acts(g[:-1], activities[:-1]) + [torch.zeros_like(activities[-1])]


def _grad_arena(net, want_energies, zero=False):
    """Both models' gradients (+ the energy pairs) as views of ONE flat buffer: one allocation per call (addresses repeat
    from step to step, so the library replays its launch plan) and ONE all-reduce under data parallelism."""
    D, n = net.dims, net.H + 1
    shapes = [(D[j + 1], D[j]) for j in range(n)] + [(D[j], D[j + 1]) for j in range(n)]
    shapes += [(D[j + 1],) for j in range(n)] if net.td_bias else []
    shapes += [(D[j],) for j in range(n)] if net.bu_bias else []
    shapes += [(2 * n,)] if want_energies else []
    lay = _flat.layout(shapes)
    flat = (torch.zeros if zero else torch.empty)(lay.total, dtype=torch.float32, device=net.dev)
    v = lay.carve(flat)
    gW, gV, k = v[:n], v[n:2 * n], 2 * n
    gbW = gbV = E = None
    if net.td_bias:
        gbW, k = v[k:k + n], k + n
    if net.bu_bias:
        gbV, k = v[k:k + n], k + n
    if want_energies:
        E = v[k]
    return flat, gW, gbW, gV, gbV, E


def _param_grads(net, activities, want_energies=False, fused_steps=False, exchange=False):
    """jpcb_bpc_param_grads on `activities` -- or, with `fused_steps`, jpcb_bpc_train_step: the descriptor's Euler steps in
    place on the (library-layout) `activities`, then the gradients at the solution.  `exchange`: data parallel -- every rank
    holds a row shard and a descriptor with the GLOBAL batch in its 1/B, so the sums over ranks of gradients and energy pairs
    are the full-batch values: one all-reduce of the flat buffer (the reference has no collective)."""
    flat, gW, gbW, gV, gbV, E = _grad_arena(net, want_energies or exchange, zero=exchange)
    ws = net.workspace()
    fn = _lib.lib.jpcb_bpc_train_step if fused_steps else _lib.lib.jpcb_bpc_param_grads
    acts = activities if fused_steps else net.pad_acts(activities)
    _lib.check(fn(*net.args(), _lib.ptr_array(acts), _lib.ptr(net.x), _lib.ptr(net.y),
                  _lib.ptr_array(gW), _lib.ptr_array(gbW) if gbW else None, _lib.ptr_array(gV),
                  _lib.ptr_array(gbV) if gbV else None, _lib.ptr(E), _lib.ptr(ws), ws.numel(), net.stream()))
    if exchange:
        import torch.distributed as dist
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if not want_energies:
        E = None

    if net.bf16:  # slice the zero-padded rows / columns off again
        T = net.true_dims
        gW = [g[:T[j + 1], :T[j]].contiguous() for j, g in enumerate(gW)]
        gV = [g[:T[j], :T[j + 1]].contiguous() for j, g in enumerate(gV)]
        gbW = [g[:T[j + 1]].contiguous() for j, g in enumerate(gbW)] if gbW else None
        gbV = [g[:T[j]].contiguous() for j, g in enumerate(gbV)] if gbV else None

    def tree(model, gw, gb):
        leaves = []
        for j in range(len(gw)):
            leaves.append(gw[j])
            if gb is not None:
                leaves.append(gb[j])
        return tree_unflatten_like(model, leaves)
    return tree(net.top_down_model, gW, gbW), tree(net.bottom_up_model, gV, gbV), E


def compute_bpc_param_grads(top_down_model, bottom_up_model, activities, y, *, x=None, skip_model=None, param_type="sp",
                            backward_energy_weight=1.0, forward_energy_weight=1.0, batch_global=None):
    """jpc/_core/_grads.py:544-612: (top_down_grads, bottom_up_grads) with the models' structure.  `batch_global` (not in the
    reference signature): the B of the 1/B normalisation when this process holds a row shard; the result is then this shard's
    PART of the gradient (no collective here -- `bpc_train_step` is the data-parallel entry point)."""
    net = _net(top_down_model, bottom_up_model, x, y, skip_model, param_type, backward_energy_weight, forward_energy_weight,
               batch_global=batch_global)
    net.check_acts(activities)
    td, bu, _ = _param_grads(net, activities)
    return td, bu


def update_bpc_activities(top_down_model, bottom_up_model, activities, optim, opt_state, output, *, input=None,
                          skip_model=None, param_type="sp", backward_energy_weight=1.0, forward_energy_weight=1.0):
    """jpc/_core/_updates.py:206-281."""
    energy, grads = compute_bpc_activity_grad(top_down_model, bottom_up_model, activities, output, x=input,
                                              skip_model=skip_model, param_type=param_type,
                                              backward_energy_weight=backward_energy_weight,
                                              forward_energy_weight=forward_energy_weight)
    activities, opt_state = update_and_apply(optim, grads, opt_state, activities)
    return {"energy": energy, "activities": activities, "grads": grads, "opt_state": opt_state}


def update_bpc_params(top_down_model, bottom_up_model, activities, top_down_optim, bottom_up_optim, top_down_opt_state,
                      bottom_up_opt_state, output, *, input=None, skip_model=None, param_type="sp",
                      backward_energy_weight=1.0, forward_energy_weight=1.0):
    """jpc/_core/_updates.py:284-372."""
    td_g, bu_g = compute_bpc_param_grads(top_down_model, bottom_up_model, activities, output, x=input, skip_model=skip_model,
                                         param_type=param_type, backward_energy_weight=backward_energy_weight,
                                         forward_energy_weight=forward_energy_weight)
    top_down_model, top_down_opt_state = update_and_apply(top_down_optim, td_g, top_down_opt_state, top_down_model)
    bottom_up_model, bottom_up_opt_state = update_and_apply(bottom_up_optim, bu_g, bottom_up_opt_state, bottom_up_model)
    return {"models": (top_down_model, bottom_up_model), "grads": (td_g, bu_g),
            "opt_states": (top_down_opt_state, bottom_up_opt_state)}


def solve_bpc_inference(top_down_model, bottom_up_model, activities, output, *, input=None, n_steps=1, dt=0.5,
                        skip_model=None, param_type="sp", backward_energy_weight=1.0, forward_energy_weight=1.0,
                        batch_global=None):
    """n_steps fused Euler steps of size dt on the bPC energy == n_steps x update_bpc_activities with
    optax.sgd(dt) (the reference's inference loop), without returning to the host in between."""
    net = _net(top_down_model, bottom_up_model, input, output, skip_model, param_type, backward_energy_weight,
               forward_energy_weight, n_steps=n_steps, dt=dt, batch_global=batch_global)
    net.check_acts(activities)
    A = net.pad_acts(activities, clone=True)
    ws = net.workspace()
    _lib.check(_lib.lib.jpcb_bpc_infer(*net.args(), _lib.ptr_array(A), _lib.ptr(net.x), _lib.ptr(net.y), _lib.ptr(ws),
                                       ws.numel(), net.stream()))
    return net.unpad_acts(A, activities)


def bpc_train_step(top_down_model, bottom_up_model, activities, output, *, input=None, n_steps=1, dt=0.5, skip_model=None,
                   param_type="sp", backward_energy_weight=1.0, forward_energy_weight=1.0, record_energies=False,
                   batch_global=None):
    """`solve_bpc_inference` followed by `compute_bpc_param_grads` at the solution, as ONE library call
    (jpcb_bpc_train_step): the inference loop and the learning step of the reference's bPC example
    (examples/bidirectional_pc.ipynb cell 8: `update_bpc_activities` x T, then the gradients `update_bpc_params` applies)
    without returning to the host and without preparing the operands twice.  Returns
    (activities, top_down_grads, bottom_up_grads, energy pairs or None).

    Data parallel (torch.distributed initialised, one process per GPU, each holding a contiguous row shard of the batch --
    `shard_rows`): samples are independent through the inference loop once every rank normalises by the GLOBAL batch
    (`batch_global`, default: the sum of the ranks' row counts), so the activities stay local and the only exchange is ONE
    all-reduce of both models' gradients and the energy pairs after the call; every rank returns the full-batch gradients."""
    from ._train import _world, global_batch
    world = _world()
    bg = int(batch_global) if batch_global is not None else global_batch(output.shape[0], output.device)
    net = _net(top_down_model, bottom_up_model, input, output, skip_model, param_type, backward_energy_weight,
               forward_energy_weight, n_steps=n_steps, dt=dt, batch_global=bg)
    net.check_acts(activities)
    A = net.pad_acts(activities, clone=True)   # functional API: inputs are never mutated
    td, bu, E = _param_grads(net, A, want_energies=record_energies, fused_steps=True, exchange=world > 1)
    return net.unpad_acts(A, activities), td, bu, E
