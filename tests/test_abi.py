"""The C-ABI library loads and exports every symbol include/jpc_b200.h declares; descriptor
validation and workspace sizing (host-only entry points) behave.  No compute call is made here:
those need a GPU and live in the `-m gpu` tests."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "jpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jpcb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from jpc_b200 import _build, _lib
    assert os.path.exists(_build.LIB)
    raw = ctypes.CDLL(_build.LIB)
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/jpc_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in jpc_b200/_lib.py"
    assert _lib.lib.jpcb_version() == 101


def _desc(dtype=0, dims=(10, 20, 20, 5), B=4):
    from jpc_b200 import _lib
    d = _lib.PcDesc()
    d.n_layers, d.batch_local, d.batch_global = len(dims) - 1, B, B
    for i, v in enumerate(dims):
        d.dims[i] = v
    for l in range(len(dims) - 1):
        d.act[l] = 0 if l == 0 else 1
        d.scalings[l] = 1.0
    d.dtype, d.loss, d.solver, d.n_steps, d.dt, d.dt_last = dtype, 0, 0, 3, 0.5, 0.5
    d.output_energy_scaling = 1.0
    return d


def test_workspace_sizing_and_validation_without_a_gpu():
    from jpc_b200 import _lib
    L = _lib.lib
    assert L.jpcb_pc_workspace_bytes(_desc()) > 0
    heun = _desc()
    heun.solver = 1
    assert L.jpcb_pc_workspace_bytes(heun) > L.jpcb_pc_workspace_bytes(_desc())  # stage state + first-stage gradient
    bad = _desc()
    bad.n_layers = 0  # (a single layer is legal: jpcb_pc_ffwd_init serves one-layer feed-forward passes)
    assert L.jpcb_pc_workspace_bytes(bad) == 0 and b"n_layers" in L.jpcb_last_error()
    bad = _desc()
    bad.act[1] = 99
    assert L.jpcb_pc_workspace_bytes(bad) == 0 and b"activation" in L.jpcb_last_error()
    # bf16 tensor-core path needs multiples of 8
    assert L.jpcb_pc_workspace_bytes(_desc(dtype=1)) == 0
    assert L.jpcb_pc_workspace_bytes(_desc(dtype=1, dims=(16, 64, 64, 8), B=8)) > 0
    # error codes map to the reference's exception types
    with pytest.raises(ValueError):
        _lib.check(_lib.EINVAL)
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.ENOTIMPL)


def test_bpc_workspace_sizing():
    from jpc_b200 import _lib
    d = _lib.BpcDesc()
    dims = (10, 32, 32, 24)
    d.td.n_layers, d.td.batch_local, d.td.batch_global = 3, 8, 8
    for i, v in enumerate(dims):
        d.td.dims[i] = v
    d.forward_energy_weight = d.backward_energy_weight = 1.0
    assert _lib.lib.jpcb_bpc_workspace_bytes(d) > 0
    d.td.solver = 1  # Heun is not on the bPC path
    assert _lib.lib.jpcb_bpc_workspace_bytes(d) == 0


def test_python_api_surface_matches_the_reference_names():
    import jpc_b200 as J
    for name in ("make_pc_step", "solve_inference", "update_pc_activities", "update_pc_params", "pc_energy_fn",
                 "init_activities_with_ffwd", "compute_pc_activity_grad", "compute_pc_param_grads", "neg_pc_activity_grad",
                 "_get_param_scalings", "make_mlp", "make_skip_model", "get_act_fn", "mse_loss", "cross_entropy_loss",
                 "compute_accuracy", "update_activities", "update_params", "bpc_energy_fn", "compute_bpc_activity_grad",
                 "compute_bpc_param_grads", "update_bpc_activities", "update_bpc_params"):
        assert callable(getattr(J, name)), name
