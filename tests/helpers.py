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
        layers.append(nn.Sequential([nn.Lambda(nn.get_act_fn(layer["act"])), lin]))
    return layers


def f32_model(omodel):
    """The oracle model with weights rounded to fp32 (what the device actually holds), in fp64."""
    return [{"W": l["W"].float().double(), "b": None if l["b"] is None else l["b"].float().double(), "act": l["act"]}
            for l in omodel]


def dev(t, device="cuda"):
    return t.to(torch.float32).contiguous().to(device)


def devs(ts, device="cuda"):
    return [dev(t, device) for t in ts]


def rel(a, b):
    return O.max_rel_err(a, b)
