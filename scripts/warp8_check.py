"""Debug aid: config-2-shaped fused train step under the two warp-level kernels vs the fp64 oracle."""
import os, sys, subprocess
sys.path.insert(0, '.')
if len(sys.argv) > 1:
    import torch
    from oracle import pc_oracle as O
    from tests.test_gpu_parity import _case, _fused_step_raw
    import jpc_b200 as J
    J.set_compute_dtype("f32")
    act = sys.argv[1]
    om, sk, x, y = _case(10, 256, 3, 784, act, 256, use_bias=True, onehot_in=True, onehot_out=False)
    ref = O.pc_train_step(om, sk, x, y, solver="heun", max_t1=25.0, dt=0.5)
    acts, arena = _fused_step_raw(om, sk, x, y, "heun", 25.0, 0.5)
    for l, (p, q) in enumerate(zip(acts, ref["activities"])):
        d = (p.double().cpu() - q).abs()
        bad = (d > 1e-5 * q.abs().max()).nonzero()
        print(act, os.environ.get("JPCB_WARP8_Q"), "act", l, "rel", O.max_rel_err(p, q), "n_bad", len(bad), bad[:6].tolist())
else:
    for act in ("relu", "tanh"):
        for q in ("0", "-1"):
            env = dict(os.environ, JPCB_WARP8_Q=q, JPCB_WARP8_KMAX="100000")
            subprocess.run([sys.executable, __file__, act], env=env)
