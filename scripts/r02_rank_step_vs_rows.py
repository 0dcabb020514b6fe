"""One rank's share of the strong-scaled config 4 (B_local rows of a 4096-row global batch) on ONE GPU, no exchange:
device time of the fused train step (with and without the weight gradients) -- what the data-parallel step has to hide
its gradient exchange behind."""
import sys, time, torch
sys.path.insert(0, '.')
import bench, jpc_b200 as J
from jpc_b200 import _lib
from jpc_b200._core import _Net, _solver_id, _time_grid
from jpc_b200._train import _Arena
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
w = dict(bench.WORKLOADS['c4'], B=B)
J.set_compute_dtype('bf16')
om, x_h, y_h = bench.make_inputs(w)
dev = torch.device('cuda')
gm = bench.to_device_model(om, dev)
x, y = x_h.to(dev), y_h.to(dev)
T, h, hl = _time_grid(w['max_t1'], w['dt'])
net = _Net(gm, None, x, y, param_type='sp', loss_id='mse', gamma=None, output_energy_scaling=None, activity_decay=0.0,
           weight_decay=0.0, spectral_penalty=0.0, solver=_solver_id(J.Euler()), n_steps=T, dt=h, dt_last=hl, batch_global=4096)
arena, acts, ws = _Arena(net), net.new_acts(), net.workspace()
A, gW = _lib.ptr_array(acts), _lib.ptr_array(arena.gW)
def step(with_gw):
    _lib.check(_lib.lib.jpcb_pc_train_step(net.desc, net.W, net.b, _lib.ptr(x), _lib.ptr(y), A, _lib.ptr(arena.loss), _lib.ptr(arena.E),
                                           gW if with_gw else None, None, _lib.ptr(ws), ws.numel(), net.stream()))
for with_gw in (True, False):
    for _ in range(5):
        step(with_gw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step(with_gw)
    e1.record()
    torch.cuda.synchronize()
    print(f'B_local={B} with_gW={with_gw}: {e0.elapsed_time(e1) / reps:.3f} ms/step', flush=True)
