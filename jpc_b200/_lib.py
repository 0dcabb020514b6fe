"""ctypes binding of the C-ABI in include/jpc_b200.h.

The CUDA library is the product: if it is missing and cannot be built, importing this module
raises -- there is no Python/CPU fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

MAX_LAYERS = 256

OK, EINVAL, ENOTIMPL, EWORKSPACE, ECUDA = 0, 1, 2, 3, 4
DTYPE_F32, DTYPE_BF16, DTYPE_F32X3 = 0, 1, 2
DTYPE_IDS = {"f32": DTYPE_F32, "bf16": DTYPE_BF16, "f32x3": DTYPE_F32X3}
LOSS_MSE, LOSS_CE = 0, 1
SOLVER_EULER, SOLVER_HEUN = 0, 1
ACT_IDS = {"linear": 0, "tanh": 1, "hard_tanh": 2, "relu": 3, "leaky_relu": 4, "gelu": 5, "selu": 6, "silu": 7}


class PcDesc(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32),
        ("batch_local", C.c_int32),
        ("batch_global", C.c_int32),
        ("dims", C.c_int32 * (MAX_LAYERS + 1)),
        ("act", C.c_int32 * MAX_LAYERS),
        ("has_bias", C.c_int32),
        ("skip", C.c_uint8 * MAX_LAYERS),
        ("scalings", C.c_float * MAX_LAYERS),
        ("dtype", C.c_int32),
        ("loss", C.c_int32),
        ("output_energy_scaling", C.c_float),
        ("activity_decay", C.c_float),
        ("solver", C.c_int32),
        ("n_steps", C.c_int32),
        ("dt", C.c_float),
        ("dt_last", C.c_float),
        ("free_input", C.c_int32),
        ("post_act", C.c_int32 * MAX_LAYERS),
    ]


class BpcDesc(C.Structure):
    _fields_ = [
        ("td", PcDesc),
        ("bu_act", C.c_int32 * MAX_LAYERS),
        ("bu_has_bias", C.c_int32),
        ("forward_energy_weight", C.c_float),
        ("backward_energy_weight", C.c_float),
    ]


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)

# every symbol include/jpc_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "jpcb_version": (C.c_int, []),
    "jpcb_last_error": (C.c_char_p, []),
    "jpcb_launch_count": (C.c_ulonglong, []),
    "jpcb_set_reproducible": (C.c_int, [C.c_int]),
    "jpcb_pc_workspace_bytes": (C.c_size_t, [C.POINTER(PcDesc)]),
    "jpcb_pc_ffwd_init": (C.c_int, [C.POINTER(PcDesc), _PP, _PP, _P, _PP, _P, _P, _P, C.c_size_t, _P]),
    "jpcb_pc_energy": (C.c_int, [C.POINTER(PcDesc), _PP, _PP, _PP, _P, _P, _P, _P, C.c_size_t, _P]),
    "jpcb_pc_activity_grad": (C.c_int, [C.POINTER(PcDesc), _PP, _PP, _PP, _P, _P, _P, _PP, _P, C.c_size_t, _P]),
    "jpcb_pc_infer": (C.c_int, [C.POINTER(PcDesc), _PP, _PP, _PP, _P, _P, _P, _P, C.c_size_t, _P]),
    "jpcb_pc_heun_trial": (C.c_int, [C.POINTER(PcDesc), _PP, _PP, _PP, _P, _P, C.c_float, C.c_float, C.c_float, _PP, _P, _P,
                                     C.c_size_t, _P]),
    "jpcb_pc_param_grads": (C.c_int, [C.POINTER(PcDesc), _PP, _PP, _PP, _P, _P, _PP, _PP, _P, _P, C.c_size_t, _P]),
    "jpcb_pc_train_step": (C.c_int, [C.POINTER(PcDesc), _PP, _PP, _P, _P, _PP, _P, _P, _PP, _PP, _P, C.c_size_t, _P]),
    "jpcb_pc_param_grads_range": (C.c_int, [C.POINTER(PcDesc), _PP, _PP, _P, C.c_int, C.c_int, _PP, _PP, _P, C.c_size_t, _P]),
    "jpcb_bpc_workspace_bytes": (C.c_size_t, [C.POINTER(BpcDesc)]),
    "jpcb_bpc_energy": (C.c_int, [C.POINTER(BpcDesc), _PP, _PP, _PP, _PP, _PP, _P, _P, _P, _P, C.c_size_t, _P]),
    "jpcb_bpc_activity_grad": (C.c_int, [C.POINTER(BpcDesc), _PP, _PP, _PP, _PP, _PP, _P, _P, _P, _PP, _P, C.c_size_t, _P]),
    "jpcb_bpc_infer": (C.c_int, [C.POINTER(BpcDesc), _PP, _PP, _PP, _PP, _PP, _P, _P, _P, C.c_size_t, _P]),
    "jpcb_bpc_param_grads": (C.c_int, [C.POINTER(BpcDesc), _PP, _PP, _PP, _PP, _PP, _P, _P, _PP, _PP, _PP, _PP, _P, _P, C.c_size_t, _P]),
    "jpcb_bpc_train_step": (C.c_int, [C.POINTER(BpcDesc), _PP, _PP, _PP, _PP, _PP, _P, _P, _PP, _PP, _PP, _PP, _P, _P, C.c_size_t, _P]),
    "jpcb_hpc_param_grads": (C.c_int, [C.POINTER(PcDesc), _PP, _PP, _PP, _PP, _PP, _PP, _P, _P, C.c_size_t, _P]),
    "jpcb_epc_workspace_bytes": (C.c_size_t, [C.POINTER(PcDesc)]),
    "jpcb_epc_grads": (C.c_int, [C.POINTER(PcDesc), _PP, _PP, _PP, _P, _P, _P, _PP, _PP, _PP, _P, C.c_size_t, _P]),
    "jpcb_pc_weight_reg_workspace_bytes": (C.c_size_t, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "jpcb_pc_weight_reg": (C.c_int, [C.c_int, _PP, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_float, C.c_float, _P, C.c_int, _PP, _P,
                                     C.c_size_t, _P]),
    "jpcb_plan_cache_configure": (C.c_int, [C.c_int, C.c_int]),
    "jpcb_plan_cache_clear": (C.c_int, []),
    "jpcb_plan_cache_stats": (None, [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "jpcb_call_counts": (C.c_int, [C.c_int, _P, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "jpcb_call": (C.c_int, [C.c_int, _P, C.c_size_t, _PP, C.c_int, _PP, C.c_int, C.c_size_t, _P]),
    "jpcb_optim_sgd": (C.c_int, [C.c_int, C.POINTER(C.c_longlong), _PP, _PP, C.c_double, _PP, _P]),
    "jpcb_optim_adam": (C.c_int, [C.c_int, C.POINTER(C.c_longlong), _PP, _PP, _PP, _PP, C.c_double, C.c_double, C.c_double,
                                  C.c_double, C.c_int, _PP, _PP, _PP, _P]),
}


def _bind(path: str) -> C.CDLL:
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def _load() -> C.CDLL:
    path = _build.LIB
    if not os.path.exists(path):
        # one attempt to build in-tree; failure is fatal (no fallback implementation exists)
        _build.build()
    try:
        return _bind(path)
    except (AttributeError, OSError):
        # stale binary from an older source tree: rebuild once, then fail loudly
        _build.build(force=True)
        return _bind(path)


lib = _load()


class JpcbError(RuntimeError):
    pass


def check(code: int) -> None:
    """Maps C-ABI error codes to the exceptions the reference raises for the same conditions."""
    if code == OK:
        return
    msg = lib.jpcb_last_error().decode("utf-8", "replace")
    if code == EINVAL:
        raise ValueError(msg)
    if code == ENOTIMPL:
        raise NotImplementedError(msg)
    raise JpcbError(f"jpc_b200 error {code}: {msg}")


def ptr_array(tensors) -> "C.Array":
    """Host array of device pointers from a list of tensors (None -> NULL)."""
    return (C.c_void_p * max(1, len(tensors)))(*[None if t is None else t.data_ptr() for t in tensors])


def ptr_array_from_ints(addrs) -> "C.Array":
    """Host array of device pointers from integer addresses."""
    return (C.c_void_p * max(1, len(addrs)))(*addrs)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


# entry ids of the positional form (include/jpc_b200.h: JPCB_CALL_*)
CALL_ENTRIES = ("pc_ffwd_init", "pc_energy", "pc_activity_grad", "pc_infer", "pc_param_grads", "pc_train_step", "hpc_param_grads",
                "epc_grads", "bpc_energy", "bpc_activity_grad", "bpc_infer", "bpc_param_grads", "bpc_train_step")


def call_counts(entry: str, desc) -> tuple:
    """(operands, results) the positional form of `entry` takes for `desc` (results include the trailing workspace)."""
    na, nr = C.c_int(0), C.c_int(0)
    check(lib.jpcb_call_counts(CALL_ENTRIES.index(entry), C.cast(C.pointer(desc), _P), C.sizeof(desc), C.byref(na), C.byref(nr)))
    return na.value, nr.value


def call(entry: str, desc, args, rets, stream) -> None:
    """jpcb_call: the positional form an XLA custom call uses -- `args` / `rets` are flat lists of device tensors, the last
    result is the workspace (a uint8 tensor)."""
    check(lib.jpcb_call(CALL_ENTRIES.index(entry), C.cast(C.pointer(desc), _P), C.sizeof(desc), ptr_array(args), len(args),
                        ptr_array(rets), len(rets), rets[-1].numel() * rets[-1].element_size(), stream))


def plan_cache_configure(max_plans: int = 16, min_calls: int = 2) -> None:
    """Launch plans of the library (include/jpc_b200.h, csrc/plan.cu): identical calls are captured into a CUDA graph the
    `min_calls`-th time they are seen and replayed from then on.  `max_plans=0` disables (every call enqueues its kernels)."""
    check(lib.jpcb_plan_cache_configure(int(max_plans), int(min_calls)))


def plan_cache_clear() -> None:
    check(lib.jpcb_plan_cache_clear())


def plan_cache_stats() -> dict:
    a, b, c = C.c_ulonglong(0), C.c_ulonglong(0), C.c_ulonglong(0)
    lib.jpcb_plan_cache_stats(C.byref(a), C.byref(b), C.byref(c))
    return {"replays": a.value, "captures": b.value, "bypassed": c.value}


def set_reproducible(on: bool = True) -> None:
    """Energies and losses by fixed-order sums instead of floating-point atomics (include/jpc_b200.h): bit-for-bit
    reproducible run to run, two small extra launches per energy evaluation.  Process-wide, default off."""
    check(lib.jpcb_set_reproducible(1 if on else 0))
