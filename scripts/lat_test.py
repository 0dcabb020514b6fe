import time, torch, sys
sys.path.insert(0, '.')
import bench, jpc_b200 as J
from jpc_b200 import _lib
from jpc_b200._core import _Net, _solver_id, _time_grid
from jpc_b200._train import _Arena
for wl in ('c1', 'c3'):
    w = bench.WORKLOADS[wl]
    J.set_compute_dtype(w['dtype'])
    om, x_h, y_h = bench.make_inputs(w)
    dev = torch.device('cuda')
    gm = bench.to_device_model(om, dev)
    gsk = J.make_skip_model(w['depth']) if w['skips'] else None
    x, y = x_h.to(dev), y_h.to(dev)
    T, h, hl = _time_grid(w['max_t1'], w['dt'])
    net = _Net(gm, gsk, x, y, param_type=w['ptype'], loss_id='mse', gamma=None, output_energy_scaling=None, activity_decay=0.0,
               weight_decay=0.0, spectral_penalty=0.0, solver=_solver_id(J.Euler()), n_steps=T, dt=h, dt_last=hl)
    arena = _Arena(net); acts = net.new_acts(); ws = net.workspace(); st = net.stream()
    A, gW = _lib.ptr_array(acts), _lib.ptr_array(arena.gW)
    def step():
        _lib.check(_lib.lib.jpcb_pc_train_step(net.desc, net.W, net.b, _lib.ptr(x), _lib.ptr(y), A, _lib.ptr(arena.loss), _lib.ptr(arena.E), gW, None, _lib.ptr(ws), ws.numel(), st))
    for _ in range(3): step()
    torch.cuda.synchronize()
    l0 = _lib.lib.jpcb_launch_count()
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    n = _lib.lib.jpcb_launch_count() - l0
    print(wl, 'launches', n, 'cpu enqueue ms', (t1-t0)*1e3, 'total ms', (t2-t0)*1e3, 'us/launch cpu', (t1-t0)*1e6/n)
