"""`PCTrainer`: the fixed-step train step of `make_pc_step` (jpc/_train.py:35-267) as ONE replayed CUDA graph.

Not part of the reference API -- an addition for the launch-bound regime (small / deep networks, BASELINE configs 1-3),
where a train step is tens to hundreds of short dependent kernel launches and the per-call host work of the functional
API (new model objects, pointer tables, output allocations) costs more than the GPU work.  The trainer owns persistent
device buffers for the parameters, the batch, the activities and the gradient arena; `jpcb_pc_train_step` plus the
in-place `jpcb_optim_sgd` update are captured once into a CUDA graph (the C-ABI only enqueues on the caller's stream
and never synchronises, so stream capture just works) and every `step()` is a copy of the batch into the static
buffers and one graph launch.  Same arithmetic, same kernels, same results as `make_pc_step`
(tests/test_gpu_parity.py::test_pc_trainer_graph_matches_make_pc_step).

Data parallel (torch.distributed initialised, world_size > 1): every rank builds its trainer collectively; the batch
arguments are the rank's row shard, the energy is normalised by the GLOBAL batch, and the NCCL all-reduce (SUM) of the
gradient arena is captured INSIDE the graph, between the gradient kernels and the update.  The captured collective runs on
a communicator of its own (`dist.new_group`), never used eagerly after the warm-up: interleaving eager collectives with
graph replays on one communicator is what hung the first attempt at this (round 1).

Restrictions: fixed-step solvers (`ConstantStepSize`), `optim.sgd` inside the graph (other optimisers run after the
graph through the optax protocol, in place)."""
from __future__ import annotations

import torch

from . import _lib
from ._core import _Net, _solver_id, _time_grid
from ._errors import _check_param_type
from ._train import _Arena, _world, global_batch
from .nn import tree_leaves, tree_unflatten_like
from .solvers import ConstantStepSize, Euler


class PCTrainer:
    def __init__(self, model, optim, *, batch_size, skip_model=None, loss_id="mse", param_type="sp", ode_solver=None,
                 max_t1=20, dt=None, stepsize_controller=None, activity_decay=0.0):
        _check_param_type(param_type)
        if not isinstance(stepsize_controller, ConstantStepSize):
            raise NotImplementedError("PCTrainer replays a fixed launch sequence: pass stepsize_controller=ConstantStepSize() and dt")
        self._world = _world()
        self._pg = None
        if self._world > 1:
            import torch.distributed as dist
            self._pg = dist.new_group(backend="nccl")  # collective: every rank constructs its trainer at the same point
        ode_solver = Euler() if ode_solver is None else ode_solver
        T, h, h_last = _time_grid(max_t1, dt)
        # persistent parameters: clones of the caller's (the caller's model is left untouched)
        leaves = [p.detach().clone().contiguous() for p in tree_leaves((model, skip_model))]
        self.model, self.skip_model = tree_unflatten_like((model, skip_model), leaves)
        self._params = leaves
        self.optim = optim
        self.opt_state = optim.init((self.model, self.skip_model))
        dev = leaves[0].device
        d_in = leaves[0].shape[1]
        d_out = [p for p in leaves if p.ndim == 2][-1].shape[0]
        self.x = torch.zeros(batch_size, d_in, dtype=torch.float32, device=dev)
        self.y = torch.zeros(batch_size, d_out, dtype=torch.float32, device=dev)
        self._net = _Net(self.model, self.skip_model, self.x, self.y, param_type=param_type, loss_id=loss_id, gamma=None,
                         output_energy_scaling=None, activity_decay=activity_decay, weight_decay=0.0, spectral_penalty=0.0,
                         solver=_solver_id(ode_solver), n_steps=T, dt=h, dt_last=h_last,
                         batch_global=global_batch(batch_size, dev) if self._world > 1 else None)
        if self._net.padded:
            raise NotImplementedError("PCTrainer needs feature dims that are multiples of 8 on the tensor-core paths "
                                      "(the functional API pads per call; a persistent padded layout is not built)")
        net = self._net
        self._arena = _Arena(net)
        self._acts = net.new_acts()
        self._ws = torch.empty(max(net.ws_bytes, 1 << 20), dtype=torch.uint8, device=dev)  # own workspace: its address is baked into the graph
        self._stream = torch.cuda.Stream(device=dev)
        self._graph = None
        self._sgd_lr = getattr(optim, "sgd_learning_rate", None)  # in-graph, in-place update for optim.sgd
        self._capture()

    def _enqueue(self):
        net, arena = self._net, self._arena
        ws = self._ws
        st = torch.cuda.current_stream(net.dev).cuda_stream
        _lib.check(_lib.lib.jpcb_pc_train_step(
            net.desc, net.W, net.b, _lib.ptr(net.x), _lib.ptr(net.y), _lib.ptr_array(self._acts), _lib.ptr(arena.loss),
            _lib.ptr(arena.E), _lib.ptr_array(arena.gW), _lib.ptr_array(arena.gb) if arena.gb else None,
            _lib.ptr(ws), ws.numel(), st))
        if self._pg is not None:  # the one exchange step (captured into the graph like the kernels around it)
            import torch.distributed as dist
            dist.all_reduce(arena.flat, op=dist.ReduceOp.SUM, group=self._pg)
        if self._sgd_lr is not None:
            import ctypes as C
            p, g = self._params, []
            for l in range(net.L):
                g.append(arena.gW[l])
                if arena.gb:
                    g.append(arena.gb[l])
            numel = (C.c_longlong * len(p))(*[t.numel() for t in p])
            _lib.check(_lib.lib.jpcb_optim_sgd(len(p), numel, _lib.ptr_array(p), _lib.ptr_array(g), float(self._sgd_lr),
                                               _lib.ptr_array(p), st))

    def _capture(self):
        with torch.cuda.stream(self._stream):
            for _ in range(3 if self._pg is not None else 2):  # warm-up outside capture (lazy module loads, function attributes, NCCL communicator set-up); undone below
                saved = [p.clone() for p in self._params]
                self._enqueue()
                for p, s in zip(self._params, saved):
                    p.copy_(s)
            self._stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                self._enqueue()
            self._graph = g
            for p, s in zip(self._params, saved):  # (capture does not execute; nothing to undo -- kept for clarity)
                p.copy_(s)
            self._stream.synchronize()

    def close(self):
        """Drops the captured graph (and with it the captured collective).  Under data parallelism call this on every rank
        before `torch.distributed.destroy_process_group()`."""
        self._graph = None

    def step(self, output, input):
        """One train step on (input, output): returns the reference's result dict (views of the trainer's static buffers:
        they are overwritten by the next step)."""
        cur = torch.cuda.current_stream(self._net.dev)
        self._stream.wait_stream(cur)
        with torch.cuda.stream(self._stream):
            self.x.copy_(input, non_blocking=True)
            self.y.copy_(output, non_blocking=True)
            self._graph.replay()
            if self._sgd_lr is None:  # any other optimiser: protocol update after the graph, written back in place
                net, arena = self._net, self._arena
                grads = []
                for l in range(net.L):
                    grads.append(arena.gW[l])
                    if arena.gb:
                        grads.append(arena.gb[l])
                fused = getattr(self.optim, "fused_apply_flat", None)
                res = fused(self._params, grads, self.opt_state) if fused is not None else None
                if res is not None:  # the library's fused update kernel (what make_pc_step runs), written back in place
                    new, self.opt_state = res
                    torch._foreach_copy_(self._params, new)
                else:
                    upd, self.opt_state = self.optim.update(grads, self.opt_state, self._params)
                    torch._foreach_add_(self._params, list(upd))
        cur.wait_stream(self._stream)
        return {"model": self.model, "skip_model": self.skip_model, "opt_state": self.opt_state, "loss": self._arena.loss,
                "energies": self._arena.E, "activities": [a.unsqueeze(0) for a in self._acts]}
