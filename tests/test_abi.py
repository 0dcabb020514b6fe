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
    assert _lib.lib.jpcb_version() == 102


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


def test_positional_call_counts_follow_the_descriptor():
    """jpcb_call_counts (host arithmetic): operands / results of the positional form for PC, free-input and bPC descriptors."""
    from jpc_b200 import _lib
    d = _desc()                      # L = 3, no bias
    assert _lib.call_counts("pc_ffwd_init", d) == (3 + 2, 3 + 1 + 1)
    assert _lib.call_counts("pc_energy", d) == (3 + 3 + 2, 2)
    d.has_bias = 1
    assert _lib.call_counts("pc_train_step", d) == (6 + 2, 3 + 2 + 6 + 1)
    assert _lib.call_counts("pc_param_grads", d) == (6 + 3 + 2, 6 + 1 + 1)
    assert _lib.call_counts("hpc_param_grads", d) == (6 + 6, 6 + 1 + 1)
    assert _lib.call_counts("epc_grads", d) == (6 + 3 + 2, 1 + 3 + 6 + 1)
    d.free_input = 1                 # the state list grows by z_0 and x disappears
    assert _lib.call_counts("pc_infer", d) == (6 + 4 + 1, 4 + 1)
    assert _lib.call_counts("pc_activity_grad", d) == (6 + 4 + 1, 1 + 4 + 1)
    b = _lib.BpcDesc()
    b.td = _desc()
    b.td.has_bias, b.bu_has_bias = 1, 0
    assert _lib.call_counts("bpc_energy", b) == (3 + 3 + 3 + 0 + 3 + 2, 2)
    assert _lib.call_counts("bpc_train_step", b) == (14, 3 + (3 + 3 + 3 + 0) + 1 + 1)
    b.td.free_input = 1
    assert _lib.call_counts("bpc_infer", b) == (13, 3 + 1)
    with pytest.raises(ValueError):  # a PC descriptor where a bPC one is expected
        _lib.call_counts("bpc_infer", d)


def test_xla_shim_registers_every_positional_entry():
    """The shim cannot be compiled here (no jaxlib headers): keep it in step with the header by text -- one handler per
    JPCB_CALL_* enumerator, and the enumerators in the order jpc_b200._lib.CALL_ENTRIES assumes."""
    from jpc_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "jpc_b200.h")).read()
    enum = re.findall(r"\b(JPCB_CALL_[A-Z_]+)\s*=\s*(\d+)", hdr)
    names = [n for n, _ in enum if n != "JPCB_CALL_N_ENTRIES"]
    assert [int(v) for n, v in enum if n != "JPCB_CALL_N_ENTRIES"] == list(range(len(names)))
    assert dict(enum)["JPCB_CALL_N_ENTRIES"] == str(len(names))
    assert [n[len("JPCB_CALL_"):].lower() for n in names] == list(_lib.CALL_ENTRIES)
    shim = open(os.path.join(ROOT, "jpc_b200", "csrc", "xla_ffi_shim.cc")).read()
    handlers = dict(re.findall(r"JPCB_XLA_HANDLER\((jpcb_xla_[a-z_]+),\s*(JPCB_CALL_[A-Z_]+)\)", shim))
    assert sorted(handlers.values()) == sorted(names)
    for sym, entry in handlers.items():
        assert sym == "jpcb_xla_" + entry[len("JPCB_CALL_"):].lower()
