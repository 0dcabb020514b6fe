"""Host-side cost of make_pc_step (python + ctypes + enqueue) for the launch-bound configs: cProfile top list."""
import cProfile, pstats, sys, time, torch
sys.path.insert(0, '.')
import bench, jpc_b200 as J
wl = sys.argv[1] if len(sys.argv) > 1 else 'c3'
w = bench.WORKLOADS[wl]
J.set_compute_dtype(w['dtype'])
om, x_h, y_h = bench.make_inputs(w)
dev = torch.device('cuda')
gm = bench.to_device_model(om, dev)
gsk = J.make_skip_model(w['depth']) if w['skips'] else None
x, y = x_h.to(dev), y_h.to(dev)
opt = J.optim.sgd(1e-3 * w['B'])
st = opt.init((gm, gsk))
solver = J.Euler() if w['solver'] == 'euler' else J.Heun()
def step(model, st):
    out = J.make_pc_step(model, opt, st, y, input=x, param_type=w['ptype'], ode_solver=solver, max_t1=w['max_t1'], dt=w['dt'],
                         stepsize_controller=J.ConstantStepSize(), skip_model=gsk)
    return out['model'], out['opt_state'], out['loss']
m = gm
for _ in range(5):
    m, st, loss = step(m, st)
float(loss)
N = 50
t0 = time.perf_counter()
for _ in range(N):
    m, st, loss = step(m, st)
t1 = time.perf_counter()
float(loss)
t2 = time.perf_counter()
print(wl, 'host ms/step (enqueue only)', (t1 - t0) / N * 1e3, 'incl. drain', (t2 - t0) / N * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    m, st, loss = step(m, st)
pr.disable()
float(loss)
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
