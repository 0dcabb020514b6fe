import sys, torch
sys.path.insert(0, '.')
from oracle import pc_oracle as O
from tests.test_gpu_parity import _case, _fused_step_raw
import jpc_b200 as J
B, D, L = 4096, 2048, 8
om, sk, x, y = _case(D, D, L, D, "tanh", B)
J.set_compute_dtype("bf16")
acts, arena = _fused_step_raw(om, sk, x, y, "euler", 10.0, 0.5)
A = [a.double().cpu() for a in acts]
errs, scal = O.closed_form_errors(om, None, A, y, x)
for l in range(L):
    g = arena.gW[l].double().cpu()
    h = x if l == 0 else O.act_fn(om[l]["act"], A[l - 1])
    ref = -(scal[l] / B) * errs[l].T @ h
    nz = (g != 0).double().mean()
    print(l, "gpu norm %.4e ref norm %.4e rel %.3e nonzero frac %.3f" % (float(g.norm()), float(ref.norm()), O.max_rel_err(g, ref), float(nz)),
          "rows nz", int((g.abs().sum(1) != 0).sum()), "cols nz", int((g.abs().sum(0) != 0).sum()))
