"""make_pc_step (jpc/_train.py:35-267) over the fused C-ABI train step.

One call = ffwd init -> loss at the init -> T-step inference -> per-layer energies at the solution
-> parameter gradients (all inside jpcb_pc_train_step, one stream, no host round trip), then --
when torch.distributed is initialised with world_size > 1 -- ONE all-reduce (SUM) of the flat
fp32 buffer [gW_0, gb_0, ..., gW_{L-1}, gb_{L-1}, E_layers, loss] over NCCL, then the optax-shaped
optimiser update, exactly where the reference leaves it to optax (jpc/_train.py:234-242).
"""
from __future__ import annotations

import torch

from . import _flat, _lib
from ._core import (_Net, _grads_pytree, _solver_id, _time_grid, init_activities_from_normal, init_activities_with_ffwd,
                    relower, solve_inference)
from ._errors import _check_param_type
from ._utils import (compute_accuracy, compute_activity_norms, compute_infer_energies, compute_param_norms,
                     get_t_max)
from .nn import tree_unflatten_like
from .optim import update_and_apply
from .solvers import ConstantStepSize, Heun, PIDController


def _world():
    """(world_size, process group available) of the data-parallel job, 1 when not distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


def _contig_strides(shape):
    st, acc = [], 1
    for n in reversed(shape):
        st.append(acc)
        acc *= n
    return tuple(reversed(st))


class _Arena:
    """One flat fp32 buffer holding every tensor that must be summed over ranks, so a train step
    needs a single collective: [gW_0, gb_0, ..., E_layers(L), loss(1)].  Segments start on
    256-byte boundaries (the kernels use 16-byte vector stores)."""

    def __init__(self, net=None, *, weight_shapes=None, bias_shapes=None, device=None):
        if net is not None:  # (library-layout shapes: zero-padded feature dims on the bf16 path)
            weight_shapes = net.w_shapes
            bias_shapes = net.b_shapes if net.has_bias else None
            device = net.dev
        L = len(weight_shapes)
        shapes = []
        for l in range(L):
            shapes.append(weight_shapes[l])
            if bias_shapes is not None:
                shapes.append(bias_shapes[l])
        shapes += [(L,), ()]
        lay = _flat.layout(shapes)  # (cached per shape list; views carved run by run: this runs every train step)
        offs = lay.offs
        self.flat = torch.empty(lay.total, dtype=torch.float32, device=device)
        self._lay = lay
        self._has_bias = bias_shapes is not None
        # [E_layers, loss] are always handed out; the per-parameter gradient views only when somebody asks (`gW` / `gb`):
        # the fused train step and the fused update work from `flat` + `offs`
        self.E = self.flat.as_strided((L,), (1,), offs[-2])
        self.loss = self.flat.as_strided((), (), offs[-1])
        self.n_layers = L
        self.per_layer = 2 if bias_shapes is not None else 1
        self.offs = offs            # start of every segment in `flat`; offs[per_layer * l] = first element of layer l
        self.tail_off = offs[-2]    # [E_layers, loss]

    def grad_ptr_arrays(self):
        """(gW, gb) host arrays of device pointers into `flat`, from the layout's offsets (no tensor views needed)."""
        base, per, L = self.flat.data_ptr(), self.per_layer, self.n_layers
        gW = _lib.ptr_array_from_ints([base + 4 * self.offs[per * l] for l in range(L)])
        gb = _lib.ptr_array_from_ints([base + 4 * self.offs[per * l + 1] for l in range(L)]) if self._has_bias else None
        return gW, gb

    def __getattr__(self, name):  # (only reached when the attribute is not set yet)
        if name in ("gW", "gb"):
            views = self._lay.carve(self.flat)
            L, per = self.n_layers, self.per_layer
            self.gW = views[0:per * L:per]
            self.gb = views[1:per * L:per] if self._has_bias else None
            return self.__dict__[name]
        raise AttributeError(name)

    def all_reduce(self):
        """The one exchange step of the data-parallel path: SUM over ranks of every rank's partial
        gradients / energies / loss (each already divided by the GLOBAL batch size)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            _reduce_slices(self, [(0, self.flat.numel())])

    def buckets(self, n_buckets: int):
        """Contiguous layer groups [(layer_lo, layer_hi, flat_lo, flat_hi)] of roughly equal size covering the whole
        arena (the tail [E_layers, loss] rides with the last group): the unit of the bucketed exchange."""
        L = self.n_layers
        n_buckets = max(1, min(int(n_buckets), L))
        starts = [self.offs[self.per_layer * l] for l in range(L)] + [self.tail_off]
        target = self.tail_off / n_buckets
        out, lo = [], 0
        for b in range(n_buckets):
            if b == n_buckets - 1:
                hi = L
            else:
                hi = lo + 1
                while hi < L - (n_buckets - 1 - b) and starts[hi] < (b + 1) * target:
                    hi += 1
            out.append((lo, hi, starts[lo], starts[hi] if hi < L else self.flat.numel()))
            lo = hi
            if lo >= L:
                break
        return out


_GRAD_REDUCE_DTYPE = "f32"


def set_grad_allreduce_dtype(name: str) -> None:
    """Payload of the data-parallel gradient exchange: "f32" (default; the N-rank result equals the single-process one
    to ~1e-7) or "bf16" (half the NVLink bytes: the weight / bias gradients travel as bf16 and are summed in bf16 -- within
    the 2e-2 of the bf16 compute path, not within the 1e-5 of the fp32 paths; energies and loss always travel as fp32)."""
    global _GRAD_REDUCE_DTYPE
    if name not in ("f32", "bf16"):
        raise ValueError('gradient all-reduce dtype must be "f32" or "bf16"')
    _GRAD_REDUCE_DTYPE = name


def _reduce_slices(arena, slices):
    """all-reduce (SUM) of `flat[lo:hi]` for every slice, on the CURRENT stream."""
    import torch.distributed as dist
    for lo, hi in slices:
        if _GRAD_REDUCE_DTYPE == "bf16" and arena.flat.is_cuda:
            w_hi = min(hi, arena.tail_off)
            if w_hi > lo:
                part = arena.flat[lo:w_hi]
                half = part.to(torch.bfloat16)
                dist.all_reduce(half, op=dist.ReduceOp.SUM)
                part.copy_(half)
            if hi > arena.tail_off:
                dist.all_reduce(arena.flat[max(lo, arena.tail_off):hi], op=dist.ReduceOp.SUM)
        else:
            dist.all_reduce(arena.flat[lo:hi], op=dist.ReduceOp.SUM)


_COMM_STREAMS: dict = {}
DP_BUCKETS = None   # layer groups of the bucketed exchange; None: chosen per call from the shard's row count


def set_dp_buckets(n) -> None:
    """Number of layer groups of the data-parallel gradient exchange (None: automatic).  Each group's weight-gradient GEMMs
    are followed by that group's all-reduce on a side stream, so the exchange of group g overlaps the GEMMs of group g + 1;
    with few rows per rank those GEMMs are too short to hide anything and one exchange of the whole arena is faster."""
    global DP_BUCKETS
    DP_BUCKETS = None if n is None else max(1, int(n))


def _auto_buckets(net) -> int:
    # weight-gradient GEMMs of a rank: 2 B_local sum(d_l d_{l+1}) FLOP.  Below ~1 k rows per rank they take tens of
    # microseconds (config 4: 30 us at 512 rows against a 134 MB exchange) and splitting only adds collective launches.
    return 4 if net.B >= 2048 else 1


def _comm_stream(dev):
    st = _COMM_STREAMS.get(dev)
    if st is None:
        st = _COMM_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return st


def dp_param_grads_and_exchange(net, arena, n_buckets=None):
    """The data-parallel tail of a train step whose jpcb_pc_train_step ran with gW = NULL: weight gradients bucket by
    bucket (one grouped GEMM launch per layer group, jpcb_pc_param_grads_range), each bucket's NCCL all-reduce enqueued on
    a side stream as soon as its launch is, so it travels over NVLink while the next bucket's GEMMs run; the caller's
    stream then waits for the last bucket.  (The energies / loss tail rides with the last bucket.)"""
    main = torch.cuda.current_stream(net.dev)
    comm = _comm_stream(net.dev)
    ws = net.workspace()
    gW_p, gb_p = _lib.ptr_array(arena.gW), (_lib.ptr_array(arena.gb) if arena.gb else None)
    arena.flat.record_stream(comm)
    if n_buckets is None:
        n_buckets = DP_BUCKETS if DP_BUCKETS is not None else _auto_buckets(net)
    for llo, lhi, flo, fhi in arena.buckets(n_buckets):
        _lib.check(_lib.lib.jpcb_pc_param_grads_range(net.desc, net.W, net.b, _lib.ptr(net.x), llo, lhi, gW_p, gb_p,
                                                      _lib.ptr(ws), ws.numel(), main.cuda_stream))
        ev = torch.cuda.Event()
        ev.record(main)
        comm.wait_event(ev)
        with torch.cuda.stream(comm):
            _reduce_slices(arena, [(flo, fhi)])
    main.wait_stream(comm)


def _fused_param_update(net, arena, optim, opt_state):
    """`optim.update` + `eqx.apply_updates` (jpc/_train.py:234-242) through the library's update kernel, fed from what this
    step already holds: the parameter addresses `_Net` gathered (and validated), the gradient addresses inside the arena,
    and ONE fresh flat buffer per output family carved into the new leaves.  (new leaves, new state), or None when the
    optimiser is not one of the two the kernel implements."""
    spec = getattr(optim, "fused_spec", None)
    if spec is None or (spec[0] == "adam" and not (isinstance(opt_state, dict) and "mu" in opt_state)):
        return None
    import ctypes as C
    L, per = net.L, arena.per_layer
    shapes = []
    for l in range(L):
        shapes.append(net.w_shapes[l])
        if per == 2:
            shapes.append(net.b_shapes[l])
    lay = _flat.layout(shapes)   # (same offsets as the front of the arena's layout)
    n = len(shapes)
    if per == 2:
        p_ptrs = [a for pair in zip(net.w_ptrs, net.b_ptrs) for a in pair]
    else:
        p_ptrs = net.w_ptrs
    g_base = arena.flat.data_ptr()
    g_ptrs = [g_base + 4 * o for o in arena.offs[:n]]
    numel = getattr(lay, "_numel_table", None)
    if numel is None:
        numel = (C.c_longlong * n)(*lay.sizes)
        try:
            lay._numel_table = numel
        except AttributeError:
            pass
    dev = net.dev
    stream = net.stream()

    def fresh():
        flat = torch.empty(lay.total, dtype=torch.float32, device=dev)
        base = flat.data_ptr()
        addrs = [base + 4 * o for o in lay.offs]
        return flat, _lib.ptr_array_from_ints(addrs), addrs
    new_flat, out_ptrs, out_addrs = fresh()
    P = _lib.ptr_array_from_ints
    if spec[0] == "sgd":
        _lib.check(_lib.lib.jpcb_optim_sgd(n, numel, P(p_ptrs), P(g_ptrs), spec[1], out_ptrs, stream))
        return lay.carve(new_flat), opt_state, out_addrs
    mu, nu = opt_state["mu"], opt_state["nu"]
    if len(mu) != n or len(nu) != n:
        return None
    count = opt_state["count"] + 1
    mo_flat, mo_ptrs, _ = fresh()
    vo_flat, vo_ptrs, _ = fresh()
    _lib.check(_lib.lib.jpcb_optim_adam(n, numel, P(p_ptrs), P(g_ptrs), _lib.ptr_array(mu), _lib.ptr_array(nu), spec[1], spec[2],
                                        spec[3], spec[4], count, out_ptrs, mo_ptrs, vo_ptrs, stream))
    return lay.carve(new_flat), {"count": count, "mu": lay.carve(mo_flat), "nu": lay.carve(vo_flat)}, out_addrs


_GLOBAL_ROWS: dict = {}


def global_batch(local_rows: int, device) -> int:
    """Rows of the GLOBAL data-parallel batch: the sum of every rank's local row count (shards need not be equal:
    `shard_rows` hands out B // N and B // N + 1 rows).  One tiny all-reduce the first time a local row count is seen;
    afterwards the shard layout is assumed stable for that local count (pass `batch_global=` to make_pc_step to skip it)."""
    world = _world()
    if world == 1:
        return int(local_rows)
    got = _GLOBAL_ROWS.get(int(local_rows))
    if got is None:
        import torch.distributed as dist
        t = torch.tensor([float(local_rows)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        got = _GLOBAL_ROWS[int(local_rows)] = int(t.item())
    return got


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous batch-row block [lo, hi) of `rank` (samples are the sharded unit)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_pc_step(model, optim, opt_state, output, *, input=None, loss_id="mse", param_type="sp", ode_solver=None,
                 max_t1=20, dt=None, stepsize_controller=None, skip_model=None, weight_decay=0.0,
                 spectral_penalty=0.0, activity_decay=0.0, key=None, layer_sizes=None, batch_size=None,
                 sigma=0.05, record_activities=False, record_energies=False, record_every=None,
                 activity_norms=False, param_norms=False, grad_norms=False, calculate_accuracy=False, batch_global=None):
    """jpc/_train.py:35-267.  Same arguments, same 13-key result dict.  `batch_global` (an addition): rows of the global
    batch when `output` / `input` are one rank's shard under torch.distributed (default: the sum of the ranks' row counts)."""
    _check_param_type(param_type)
    if input is None and any(arg is None for arg in (key, layer_sizes, batch_size)):
        raise ValueError("""
            If there is no input (i.e. `x = None`), then unsupervised training 
            is assumed, and `key`, `layer_sizes` and `batch_size` must be 
            passed for random initialisation of activities.
        """)
    if loss_id not in ("mse", "ce"):
        raise ValueError("`mse` and `ce` are the only valid losses.")
    if record_energies:
        record_activities = True
    recording = record_activities or record_every is not None
    ode_solver = Heun() if ode_solver is None else ode_solver
    stepsize_controller = PIDController(rtol=1e-3, atol=1e-3) if stepsize_controller is None else stepsize_controller
    world = _world()
    if recording and world > 1:
        raise NotImplementedError("recorded trajectories are per-process: not available under data parallelism")
    bg = int(batch_global) if batch_global is not None else global_batch(output.shape[0], output.device)
    common = dict(param_type=param_type, loss_id=loss_id, gamma=None, output_energy_scaling=None, activity_decay=activity_decay,
                  weight_decay=weight_decay, spectral_penalty=spectral_penalty, batch_global=bg)
    t_max, sol, rec_energies, exchanged, lead_acts = 0, None, None, False, False
    if input is None and not isinstance(stepsize_controller, (PIDController, ConstantStepSize)):
        raise NotImplementedError(f"unsupported step size controller {stepsize_controller!r}")
    if input is None:
        # unsupervised PC (jpc/_train.py:143-149,159-161): random-normal activities incl. z_0, no loss, then the same
        # solve -> energies -> parameter gradients; every activity list has L + 1 entries
        if batch_size != output.shape[0]:
            raise ValueError(f"batch_size={batch_size} does not match the target's {output.shape[0]} rows")
        net = _Net(model, skip_model, None, output, **common)
        arena = _Arena(net)
        arena.loss.zero_()
        arena.loss = None  # the reference returns loss=None without an input
        acts0 = init_activities_from_normal(key, layer_sizes, "unsupervised", batch_size, sigma, device=output.device)
        sol = solve_inference((model, skip_model), acts0, output, input=None, loss_id=loss_id, param_type=param_type,
                              solver=ode_solver, max_t1=max_t1, dt=dt, stepsize_controller=stepsize_controller,
                              weight_decay=weight_decay, spectral_penalty=spectral_penalty, activity_decay=activity_decay,
                              batch_global=bg, record_iters=record_activities,
                              record_every=record_every)
        t_max = get_t_max(sol) if record_activities else 0
        acts = net.in_acts([a[t_max].contiguous() for a in sol])
        if record_energies:
            rec_energies = compute_infer_energies((model, skip_model), sol, t_max, output, x=None, loss=loss_id,
                                                  param_type=param_type, weight_decay=weight_decay,
                                                  spectral_penalty=spectral_penalty, activity_decay=activity_decay)
        ws = net.workspace()
        _lib.check(_lib.lib.jpcb_pc_param_grads(net.desc, net.W, net.b, _lib.ptr_array(acts), None, _lib.ptr(net.y),
                                                _lib.ptr_array(arena.gW), _lib.ptr_array(arena.gb) if arena.gb else None,
                                                _lib.ptr(arena.E), _lib.ptr(ws), ws.numel(), net.stream()))
    elif isinstance(stepsize_controller, ConstantStepSize) and not recording:
        # fixed steps: the whole step up to the optimiser is ONE C-ABI call
        T, h, h_last = _time_grid(max_t1, dt)
        if T > 4096:
            raise RuntimeError("The maximum number of solver steps (4096) was reached.")
        net = _Net(model, skip_model, input, output, solver=_solver_id(ode_solver), n_steps=T, dt=h, dt_last=h_last, **common)
        if tuple(output.shape) != (net.B, net.true_dims[-1]):
            raise ValueError(f"output has shape {tuple(output.shape)}, expected {(net.B, net.true_dims[-1])}")
        arena = _Arena(net)
        # (the result's leaves carry the reference's leading time axis (1, B, d): made as such when nothing is sliced off)
        acts = net.new_acts(lead=() if net.padded else (1,))
        lead_acts = not net.padded
        ws = net.workspace()
        bucketed = world > 1 and net.dev.type == "cuda"
        gW_p, gb_p = (None, None) if bucketed else arena.grad_ptr_arrays()
        _lib.check(_lib.lib.jpcb_pc_train_step(
            net.desc, net.W, net.b, _lib.ptr(net.x), _lib.ptr(net.y), _lib.ptr_array(acts), _lib.ptr(arena.loss),
            _lib.ptr(arena.E), gW_p, gb_p, _lib.ptr(ws), ws.numel(), net.stream()))
        if bucketed:   # weight gradients + their exchange, bucket by bucket (the exchange overlaps the remaining GEMMs)
            dp_param_grads_and_exchange(net, arena)
            exchanged = True
    elif isinstance(stepsize_controller, (PIDController, ConstantStepSize)):
        # adaptive steps (the reference's default) or a recorded trajectory: init + loss, the solve (controller-driven
        # trial steps / one step per call when rows are saved), then energies + gradients at the selected row
        net = _Net(model, skip_model, input, output, **common)
        if tuple(output.shape) != (net.B, net.true_dims[-1]):
            raise ValueError(f"output has shape {tuple(output.shape)}, expected {(net.B, net.true_dims[-1])}")
        arena = _Arena(net)
        acts0 = net.new_acts()
        ws = net.workspace()
        _lib.check(_lib.lib.jpcb_pc_ffwd_init(net.desc, net.W, net.b, _lib.ptr(net.x), _lib.ptr_array(acts0), _lib.ptr(net.y),
                                              _lib.ptr(arena.loss), _lib.ptr(ws), ws.numel(), net.stream()))
        sol = solve_inference((model, skip_model), net.out_acts(acts0), output, input=input, loss_id=loss_id, param_type=param_type,
                              solver=ode_solver, max_t1=max_t1, dt=dt, stepsize_controller=stepsize_controller,
                              weight_decay=weight_decay, spectral_penalty=spectral_penalty, activity_decay=activity_decay,
                              batch_global=bg, record_iters=record_activities,
                              record_every=record_every)
        # jpc/_train.py:183,199-201,223-225: row t_max of a recorded trajectory, row 0 otherwise (with `record_every`
        # alone row 0 is the state at ts[0] = 0, i.e. the initial activities -- the reference does the same)
        t_max = get_t_max(sol) if record_activities else 0
        acts = net.in_acts([a[t_max].contiguous() for a in sol])
        if record_energies:
            rec_energies = compute_infer_energies((model, skip_model), sol, t_max, output, x=input, loss=loss_id,
                                                  param_type=param_type, weight_decay=weight_decay,
                                                  spectral_penalty=spectral_penalty, activity_decay=activity_decay)
        ws = net.workspace()
        _lib.check(_lib.lib.jpcb_pc_param_grads(net.desc, net.W, net.b, _lib.ptr_array(acts), _lib.ptr(net.x), _lib.ptr(net.y),
                                                _lib.ptr_array(arena.gW), _lib.ptr_array(arena.gb) if arena.gb else None,
                                                _lib.ptr(arena.E), _lib.ptr(ws), ws.numel(), net.stream()))
    else:
        raise NotImplementedError(f"unsupported step size controller {stepsize_controller!r}")
    if not exchanged:
        arena.all_reduce()  # the one exchange step of the data-parallel path (no-op when not distributed)
    acts = net.out_acts(acts)
    gW = gb = None
    if net.reg is not None or grad_norms or net.padded:
        # user-shaped gradient tensors are needed: views of the arena (slices of it on a padded tensor-core layout).  The
        # weight regularisers of bare Linear layers are added here -- once, after the exchange -- on the layers' own weights
        gW, gb = net.out_grads(arena.gW, arena.gb)
        net.weight_reg(gW=gW)
    p_norms = compute_param_norms((model, skip_model)) if param_norms else (None, None)
    g_norms = compute_param_norms(_grads_pytree(net, gW, gb)) if grad_norms else (None, None)
    a_norms = compute_activity_norms(acts) if activity_norms else None
    # the update: parameter and gradient leaves are already at hand as flat lists in tree order (per layer: weight, bias;
    # the skip model has no array leaves), so the fused optimiser kernel is fed without walking the pytrees
    fused = getattr(optim, "fused_apply_flat", None)
    res, addrs = None, None
    if not net.padded:
        res = _fused_param_update(net, arena, optim, opt_state)
        if res is not None:
            res, addrs = res[:2], res[2]
    if res is None and gW is None:
        gW, gb = net.out_grads(arena.gW, arena.gb)   # (gradient views are only carved on this, slower, path)
    if res is None and fused is not None:
        p_leaves, g_leaves = [], []
        for l, lin in enumerate(net.linears):
            p_leaves.append(lin.weight)
            g_leaves.append(gW[l])
            if lin.bias is not None:
                p_leaves.append(lin.bias)
                g_leaves.append(gb[l])
        res = fused(p_leaves, g_leaves, opt_state)
    if res is not None:
        new_leaves, opt_state = res
        model, skip_model = tree_unflatten_like((model, skip_model), new_leaves)
        if addrs is not None:
            relower(net, model, new_leaves, addrs)   # (the next step starts from what this one already knows about its model)
    else:
        (model, skip_model), opt_state = update_and_apply(optim, _grads_pytree(net, gW, gb), opt_state, (model, skip_model))
    if calculate_accuracy and input is None:
        raise ValueError("calculate_accuracy needs an input (the accuracy is that of the feed-forward pass)")
    acc = compute_accuracy(output, init_activities_with_ffwd(model, input, skip_model=skip_model,
                                                             param_type=param_type)[-1]) if calculate_accuracy else None
    return {
        "model": model,
        "skip_model": skip_model,
        "opt_state": opt_state,
        "loss": arena.loss,
        "acc": acc,
        "activities": sol if recording else (acts if lead_acts else [a.unsqueeze(0) for a in acts]),
        "t_max": t_max,
        "energies": arena.E if rec_energies is None else rec_energies,
        "activity_norms": a_norms,
        "model_param_norms": p_norms[0],
        "skip_model_param_norms": p_norms[1],
        "model_grad_norms": g_norms[0],
        "skip_model_grad_norms": g_norms[1],
    }
