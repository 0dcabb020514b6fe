"""Shared test helpers: build the same network for the oracle (CPU, fp64) and for jpc_b200 (CUDA,
fp32) from one seeded weight draw; SURVEY 8d synthetic-input conventions."""
import torch

from oracle import pc_oracle as O


def data(B, d_in, d_out, seed=1, dtype=torch.float64, onehot_out=True, onehot_in=False):
    g = torch.Generator().manual_seed(seed)
    if onehot_in:
        x = torch.nn.functional.one_hot(torch.randint(0, d_in, (B,), generator=g), d_in).to(dtype)
    else:
        x = torch.randn(B, d_in, generator=g, dtype=torch.float64).to(dtype)
    if onehot_out:
        y = torch.nn.functional.one_hot(torch.randint(0, d_out, (B,), generator=g), d_out).to(dtype)
    else:
        y = torch.randn(B, d_out, generator=g, dtype=torch.float64).to(dtype)
    return x, y


def perturb(acts, seed=2, scale=0.3):
    g = torch.Generator().manual_seed(seed)
    return [a + scale * torch.randn(a.shape, generator=g, dtype=torch.float64).to(a.dtype) for a in acts]


def to_gpu_model(omodel, device="cuda"):
    """oracle layer dicts -> jpc_b200.nn model (fp32 weights on the device)."""
    from jpc_b200 import nn
    layers = []
    for layer in omodel:
        lin = nn.Linear.__new__(nn.Linear)
        lin.weight = layer["W"].to(torch.float32).contiguous().to(device)
        lin.bias = None if layer["b"] is None else layer["b"].to(torch.float32).contiguous().to(device)
        lin.out_features, lin.in_features = layer["W"].shape
        lin.use_bias = layer["b"] is not None
        post = layer.get("post")
        if post and post != "linear":   # Form B: Sequential([Linear, Lambda(act)]) -- or [Lambda, Linear, Lambda] with an input act too
            pre = [] if layer["act"] == "linear" else [nn.Lambda(nn.get_act_fn(layer["act"]))]
            layers.append(nn.Sequential(pre + [lin, nn.Lambda(nn.get_act_fn(post))]))
        elif layer.get("bare"):
            layers.append(lin)
        else:
            layers.append(nn.Sequential([nn.Lambda(nn.get_act_fn(layer["act"])), lin]))
    return layers


def f32_model(omodel):
    """The oracle model with weights rounded to fp32 (what the device actually holds), in fp64."""
    return [{"W": l["W"].float().double(), "b": None if l["b"] is None else l["b"].float().double(), "act": l["act"],
             "post": l.get("post"), "bare": l.get("bare")} for l in omodel]


def dev(t, device="cuda"):
    return t.to(torch.float32).contiguous().to(device)


def devs(ts, device="cuda"):
    return [dev(t, device) for t in ts]


def rel(a, b):
    return O.max_rel_err(a, b)


# ---- per-tensor parity gate with a recorded, bounded group-relative arm ------------------------------------
GROUP_ARM = []     # (tag, index, error / own max, error / group max) of tensors that passed only relative to their group
NOISE_ARM = []     # (tag, index, error / own max, error / noise floor) of tensors that passed only inside their stated noise floor
N_GATED = [0]


def gate_all(got, want, tol, tag, zero_abs=1e-12, noise=None):
    """Every tensor of `got` against the oracle's `want`: max|a-b| / max|b| <= tol PER TENSOR; a tensor that is exactly
    zero in the oracle must be within `zero_abs`.  A tensor that misses its own gate but whose error is within `tol` of
    the LARGEST tensor of the list (e.g. a gradient orders of magnitude smaller than its neighbours, where fp32 rounding
    of the shared intermediate dominates) still passes, but is recorded in GROUP_ARM; test_zz_group_arm_budget prints
    those and bounds their number, so this arm cannot silently swallow a layer.  `noise` replaces the group arm by a
    per-tensor absolute floor stated by the caller (recorded in NOISE_ARM the same way)."""
    got = [g.detach().double().cpu() for g in got]
    want = [w.detach().double().cpu() for w in want]
    assert len(got) == len(want), f"{tag}: {len(got)} tensors, expected {len(want)}"
    gmax = max(float(w.abs().max()) for w in want)
    for i, (p, q) in enumerate(zip(got, want)):
        N_GATED[0] += 1
        err, qmax = float((p - q).abs().max()), float(q.abs().max())
        if qmax == 0.0:
            assert err <= zero_abs, f"{tag}[{i}]: oracle tensor is exactly zero, got max |.| = {err:.3e}"
            continue
        if err <= tol * qmax:
            continue
        if noise is not None:
            # `noise[i]`: an absolute error floor for tensor i that the CALLER derived from the arithmetic (e.g. the
            # effect of rounding the GEMM operands to bf16, measured on the host for the same data)
            assert err <= tol * qmax + float(noise[i]), \
                f"{tag}[{i}]: error {err:.3e}, own max {qmax:.3e}, tol {tol}, noise floor {float(noise[i]):.3e}"
            NOISE_ARM.append((tag, i, err / qmax, err / max(float(noise[i]), 1e-300)))
            continue
        assert err <= tol * gmax, f"{tag}[{i}]: error {err:.3e}, own max {qmax:.3e}, group max {gmax:.3e}, tol {tol}"
        GROUP_ARM.append((tag, i, err / qmax, err / gmax))
