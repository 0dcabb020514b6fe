"""Debug aid: config-5 bPC inference, fp32 CUDA-core vs f32x3 tensor-core path against the fp64 oracle."""
import sys, torch
sys.path.insert(0, '.')
from oracle import pc_oracle as O
from tests.test_gpu_parity import _bpc_case, _gpu_skip
from tests.helpers import dev, devs, to_gpu_model
import jpc_b200 as J
td, bu, sk, x, y, acts = _bpc_case(10, 1024, 6, 784, "relu", 1024)
T, dt = int(sys.argv[1]) if len(sys.argv) > 1 else 30, 0.5
a = [t.clone() for t in acts]
for _ in range(T):
    _, g = O.compute_bpc_activity_grad(td, bu, a, y, x=x, skips=sk)
    a = [p - dt * q for p, q in zip(a, g)]
for dtp in ("f32", "f32x3"):
    J.set_compute_dtype(dtp)
    ga = J.solve_bpc_inference(to_gpu_model(td), to_gpu_model(bu), devs(acts), dev(y), input=dev(x), n_steps=T, dt=dt)
    for l, (p, q) in enumerate(zip(ga[:-1], a[:-1])):
        d = (p.double().cpu() - q).abs()
        lim = 1e-5 * float(q.abs().max())
        print(dtp, "act", l, "rel %.3e" % O.max_rel_err(p, q), "n_bad", int((d > lim).sum()), "of", d.numel(), "median err %.2e" % float(d.median()))
    # one gradient evaluation (no time stepping): pure GEMM accuracy
    F, g = O.compute_bpc_activity_grad(td, bu, acts, y, x=x, skips=sk)
    gF, gg = J.compute_bpc_activity_grad(to_gpu_model(td), to_gpu_model(bu), devs(acts), dev(y), x=dev(x))
    gmax = max(float(t.abs().max()) for t in g)
    print(dtp, "single grad: max err / gmax = %.3e" % max(float((p.double().cpu() - q).abs().max()) / gmax for p, q in zip(gg[:-1], g[:-1])))
