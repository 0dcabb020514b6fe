import sys, torch
sys.path.insert(0, '.')
from oracle import pc_oracle as O
from tests.test_gpu_parity import _case, _gpu_skip
from tests.helpers import dev, devs, to_gpu_model
import jpc_b200 as J
for name, args, kw in (("skips+pad", (50, 96, 4, 10, "relu", 32), dict(skips=True, onehot_out=False)),
                       ("pad only", (50, 96, 4, 10, "relu", 32), dict(skips=False, onehot_out=False)),
                       ("skips only", (48, 96, 4, 16, "relu", 32), dict(skips=True, onehot_out=False))):
    om, sk, x, y = _case(*args, **kw)
    g = torch.Generator().manual_seed(11)
    dims = [l["W"].shape[0] for l in om]
    errs = [(0.3 * torch.randn(x.shape[0], w, generator=g, dtype=torch.float64)).float().double() for w in dims]
    F, ge = O.compute_epc_error_grad(om, sk, errs, y, x=x)
    J.set_compute_dtype("bf16")
    gF, gge = J.compute_epc_error_grad((to_gpu_model(om), _gpu_skip(sk)), devs(errs), dev(y), x=dev(x))
    print(name, "F", float(gF), float(F))
    for l, (p, q) in enumerate(zip(gge, ge)):
        d = (p.double().cpu() - q).abs()
        i = int(d.argmax())
        print("  layer", l, "max err", float(d.max()), "at", divmod(i, q.shape[1]), "qmax", float(q.abs().max()), "n>1e-3:", int((d > 1e-3).sum()))
