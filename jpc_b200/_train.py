"""make_pc_step (jpc/_train.py:35-267) over the fused C-ABI train step.

One call = ffwd init -> loss at the init -> T-step inference -> per-layer energies at the solution
-> parameter gradients (all inside jpcb_pc_train_step, one stream, no host round trip), then --
when torch.distributed is initialised with world_size > 1 -- ONE all-reduce (SUM) of the flat
fp32 buffer [gW_0, gb_0, ..., gW_{L-1}, gb_{L-1}, E_layers, loss] over NCCL, then the optax-shaped
optimiser update, exactly where the reference leaves it to optax (jpc/_train.py:234-242).
"""
from __future__ import annotations

import torch

from . import _lib
from ._core import (_Net, _grads_pytree, _solver_id, _time_grid, init_activities_from_normal, init_activities_with_ffwd,
                    solve_inference)
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
        sizes = [int(torch.Size(sh).numel()) for sh in shapes]
        offs, off = [], 0
        for s in sizes:
            offs.append(off)
            off += (s + 63) // 64 * 64
        self.flat = torch.empty(off, dtype=torch.float32, device=device)
        # one as_strided per view (contiguous strides of each shape at its offset): this runs every train step
        views = [self.flat.as_strided(sh, _contig_strides(sh), o) for o, sh in zip(offs, shapes)]
        self.gW, self.gb = [], ([] if bias_shapes is not None else None)
        i = 0
        for l in range(L):
            self.gW.append(views[i]); i += 1
            if bias_shapes is not None:
                self.gb.append(views[i]); i += 1
        self.E = views[i]
        self.loss = views[i + 1]

    def all_reduce(self):
        """The one exchange step of the data-parallel path: SUM over ranks of every rank's partial
        gradients / energies / loss (each already divided by the GLOBAL batch size)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous batch-row block [lo, hi) of `rank` (samples are the sharded unit)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_pc_step(model, optim, opt_state, output, *, input=None, loss_id="mse", param_type="sp", ode_solver=None,
                 max_t1=20, dt=None, stepsize_controller=None, skip_model=None, weight_decay=0.0,
                 spectral_penalty=0.0, activity_decay=0.0, key=None, layer_sizes=None, batch_size=None,
                 sigma=0.05, record_activities=False, record_energies=False, record_every=None,
                 activity_norms=False, param_norms=False, grad_norms=False, calculate_accuracy=False):
    """jpc/_train.py:35-267.  Same arguments, same 13-key result dict."""
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
    common = dict(param_type=param_type, loss_id=loss_id, gamma=None, output_energy_scaling=None, activity_decay=activity_decay,
                  weight_decay=weight_decay, spectral_penalty=spectral_penalty, batch_global=output.shape[0] * world)
    t_max, sol, rec_energies = 0, None, None
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
                              batch_global=output.shape[0] * world, record_iters=record_activities,
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
        acts = net.new_acts()
        ws = net.workspace()
        _lib.check(_lib.lib.jpcb_pc_train_step(
            net.desc, net.W, net.b, _lib.ptr(net.x), _lib.ptr(net.y), _lib.ptr_array(acts), _lib.ptr(arena.loss),
            _lib.ptr(arena.E), _lib.ptr_array(arena.gW), _lib.ptr_array(arena.gb) if arena.gb else None,
            _lib.ptr(ws), ws.numel(), net.stream()))
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
                              batch_global=input.shape[0] * world, record_iters=record_activities,
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
    arena.all_reduce()  # the one exchange step of the data-parallel path (no-op when not distributed)
    acts = net.out_acts(acts)
    gW, gb = net.out_grads(arena.gW, arena.gb)
    p_norms = compute_param_norms((model, skip_model)) if param_norms else (None, None)
    g_norms = compute_param_norms(_grads_pytree(net, gW, gb)) if grad_norms else (None, None)
    a_norms = compute_activity_norms(acts) if activity_norms else None
    # the update: parameter and gradient leaves are already at hand as flat lists in tree order (per layer: weight, bias;
    # the skip model has no array leaves), so the fused optimiser kernel is fed without walking the pytrees
    fused = getattr(optim, "fused_apply_flat", None)
    res = None
    if fused is not None:
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
        "activities": sol if recording else [a.unsqueeze(0) for a in acts],
        "t_max": t_max,
        "energies": arena.E if rec_energies is None else rec_energies,
        "activity_norms": a_norms,
        "model_param_norms": p_norms[0],
        "skip_model_param_norms": p_norms[1],
        "model_grad_norms": g_norms[0],
        "skip_model_grad_norms": g_norms[1],
    }
