"""Is the fused train step capturable in a CUDA graph, and what does replay cost for the launch-bound configs?
(The C-ABI enqueues on the caller's stream and never synchronises, so stream capture should just work.)"""
import sys, time, torch
sys.path.insert(0, '.')
import bench, jpc_b200 as J
from jpc_b200 import _lib
from jpc_b200._core import _Net, _solver_id, _time_grid
from jpc_b200._train import _Arena
for wl in (sys.argv[1:] or ['c1', 'c3', 'c2']):
    w = bench.WORKLOADS[wl]
    J.set_compute_dtype(w['dtype'])
    om, x_h, y_h = bench.make_inputs(w)
    dev = torch.device('cuda')
    gm = bench.to_device_model(om, dev)
    gsk = J.make_skip_model(w['depth']) if w['skips'] else None
    x, y = x_h.to(dev), y_h.to(dev)
    T, h, hl = _time_grid(w['max_t1'], w['dt'])
    solver = J.Euler() if w['solver'] == 'euler' else J.Heun()
    net = _Net(gm, gsk, x, y, param_type=w['ptype'], loss_id='mse', gamma=None, output_energy_scaling=None, activity_decay=0.0,
               weight_decay=0.0, spectral_penalty=0.0, solver=_solver_id(solver), n_steps=T, dt=h, dt_last=hl)
    arena = _Arena(net); acts = net.new_acts()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        ws = net.workspace()
        A, gW = _lib.ptr_array(acts), _lib.ptr_array(arena.gW)
        gb = _lib.ptr_array(arena.gb) if arena.gb else None
        def step():
            _lib.check(_lib.lib.jpcb_pc_train_step(net.desc, net.W, net.b, _lib.ptr(x), _lib.ptr(y), A, _lib.ptr(arena.loss), _lib.ptr(arena.E),
                                                   gW, gb, _lib.ptr(ws), ws.numel(), torch.cuda.current_stream().cuda_stream))
        for _ in range(3): step()
        side.synchronize()
        ref_loss, ref_g = float(arena.loss), arena.gW[0].clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): step()
        e1.record(); side.synchronize()
        plain = e0.elapsed_time(e1) / 50
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, stream=side):
                step()
        except Exception as ex:
            print(wl, 'capture FAILED:', repr(ex)[:200]); continue
        arena.gW[0].zero_()
        g.replay(); side.synchronize()
        ok = abs(float(arena.loss) - ref_loss) < 1e-6 * abs(ref_loss) and torch.equal(arena.gW[0], ref_g)
        e0.record()
        for _ in range(50): g.replay()
        e1.record(); side.synchronize()
        print(f"{wl}: stream launches {plain:.3f} ms/step, graph replay {e0.elapsed_time(e1) / 50:.3f} ms/step, replay reproduces the step: {ok}")
