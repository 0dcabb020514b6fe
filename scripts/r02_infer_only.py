"""Only the roofline launch of bench.py: T-step jpcb_pc_infer on config 4 from the feed-forward activities (dense schedule).
For ncu: `ncu --set full -k regex:tc_staged -s 2 -c 1 ... python scripts/r02_infer_only.py`."""
import sys, torch
sys.path.insert(0, '.')
import bench, jpc_b200 as J
from jpc_b200 import _lib
from jpc_b200._core import _Net, _solver_id, _time_grid
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'c4']
J.set_compute_dtype(w['dtype'])
om, x_h, y_h = bench.make_inputs(w)
dev = torch.device('cuda')
gm = bench.to_device_model(om, dev)
x, y = x_h.to(dev), y_h.to(dev)
T, h, hl = _time_grid(w['max_t1'], w['dt'])
net = _Net(gm, None, x, y, param_type=w['ptype'], loss_id='mse', gamma=None, output_energy_scaling=None, activity_decay=0.0,
           weight_decay=0.0, spectral_penalty=0.0, solver=_solver_id(J.Euler()), n_steps=T, dt=h, dt_last=hl)
acts = J.init_activities_with_ffwd(gm, x)
ws = net.workspace()
A = _lib.ptr_array(acts)
for _ in range(4):
    _lib.check(_lib.lib.jpcb_pc_infer(net.desc, net.W, net.b, A, _lib.ptr(x), _lib.ptr(y), None, _lib.ptr(ws), ws.numel(), net.stream()))
torch.cuda.synchronize()
print("ok", float(acts[0][0, 0]))
