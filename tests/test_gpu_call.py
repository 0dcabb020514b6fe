"""The positional form of the C-ABI (jpcb_call, csrc/call.cu) against the structured entry points, entry by entry.

jpcb_call is what the XLA custom-call shim (csrc/xla_ffi_shim.cc -- not compilable here, no jaxlib) forwards to: flat operand
and result lists, the descriptor as bytes.  Each test builds one descriptor with the host mirror, runs the structured entry
point and the positional one into separate result buffers and compares (same kernels, same order: equal up to the order of
the floating-point atomics of energy sums)."""
import ctypes as C

import pytest
import torch

from oracle import pc_oracle as O
from tests.helpers import data, dev, devs, perturb, to_gpu_model

pytestmark = pytest.mark.gpu

B = 8


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30))


def _same(got, want, tol=2e-6):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape and _rel(g, w) <= tol, (_rel(g, w), tuple(g.shape))


def _empty(shapes):
    return [torch.full(tuple(s), float("nan"), device="cuda") for s in shapes]


def _ws(nbytes):
    return torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device="cuda")


def _pc(free=False, dtype="f32", use_bias=True):
    import jpc_b200 as J
    from jpc_b200 import _core, _lib
    J.set_compute_dtype(dtype)
    om = O.make_mlp(0, 16, 24, 3, 8, "tanh", use_bias=use_bias)
    gm = to_gpu_model(om)
    x, y = data(B, 16, 8)
    net = _core._Net(gm, None, None if free else dev(x), dev(y), param_type="sp", loss_id="mse", gamma=None,
                     output_energy_scaling=None, activity_decay=0.0, weight_decay=0.0, spectral_penalty=0.0,
                     solver=_lib.SOLVER_EULER, n_steps=4, dt=0.5, dt_last=0.25)
    a0 = perturb(O.init_activities_with_ffwd(om, x))
    if free:
        a0 = [x + 0.1] + a0
    return net, devs(a0)


def _params(net):
    Ws, bs = net._keep
    return list(Ws), (list(bs) if net.has_bias else [])


@pytest.mark.parametrize("dtype,use_bias", [("f32", True), ("f32", False), ("bf16", True)])
def test_pc_entries_positional_equals_structured(dtype, use_bias):
    from jpc_b200 import _lib
    L_ = _lib.lib
    P, p = _lib.ptr_array, _lib.ptr
    net, acts = _pc(dtype=dtype, use_bias=use_bias)
    d, L = net.desc, net.L
    Ws, bs = _params(net)
    bp = P(bs) if bs else None
    ws = _ws(max(L_.jpcb_pc_workspace_bytes(d), L_.jpcb_epc_workspace_bytes(d)))
    st = net.stream() if hasattr(net, "stream") else torch.cuda.current_stream().cuda_stream
    wsh, bsh = [w.shape for w in Ws], [b.shape for b in bs]
    ash = [a.shape for a in acts]
    x, y = net.x, net.y
    assert _lib.call_counts("pc_train_step", d) == (L + len(bs) + 2, L + 2 + L + len(bs) + 1)

    # ffwd init
    A1, l1 = _empty(ash), _empty([()])[0]
    _lib.check(L_.jpcb_pc_ffwd_init(d, P(Ws), bp, p(x), P(A1), p(y), p(l1), p(ws), ws.numel(), st))
    A2, l2 = _empty(ash), _empty([()])[0]
    _lib.call("pc_ffwd_init", d, Ws + bs + [x, y], A2 + [l2, ws], st)
    _same(A2 + [l2], A1 + [l1])
    # energy
    E1, E2 = _empty([(L,)])[0], _empty([(L,)])[0]
    _lib.check(L_.jpcb_pc_energy(d, P(Ws), bp, P(acts), p(x), p(y), p(E1), p(ws), ws.numel(), st))
    _lib.call("pc_energy", d, Ws + bs + acts + [x, y], [E2, ws], st)
    _same([E2], [E1])
    # activity gradient
    F1, g1 = _empty([()])[0], _empty(ash)
    _lib.check(L_.jpcb_pc_activity_grad(d, P(Ws), bp, P(acts), p(x), p(y), p(F1), P(g1), p(ws), ws.numel(), st))
    F2, g2 = _empty([()])[0], _empty(ash)
    _lib.call("pc_activity_grad", d, Ws + bs + acts + [x, y], [F2] + g2 + [ws], st)
    _same([F2] + g2, [F1] + g1)
    # inference: fresh result buffers (the operands must stay untouched) and donated operands (results ARE the operands)
    S1 = [a.clone() for a in acts]
    _lib.check(L_.jpcb_pc_infer(d, P(Ws), bp, P(S1), p(x), p(y), None, p(ws), ws.numel(), st))
    keep = [a.clone() for a in acts]
    S2 = _empty(ash)
    _lib.call("pc_infer", d, Ws + bs + acts + [x, y], S2 + [ws], st)
    _same(S2, S1)
    assert all(torch.equal(a, k) for a, k in zip(acts, keep))
    S3 = [a.clone() for a in acts]
    _lib.call("pc_infer", d, Ws + bs + S3 + [x, y], S3 + [ws], st)
    _same(S3, S1)
    assert _rel(S1[0], acts[0]) > 1e-4   # (the solve did move the state)
    # parameter gradients
    gW1, gb1, E1 = _empty(wsh), _empty(bsh), _empty([(L,)])[0]
    _lib.check(L_.jpcb_pc_param_grads(d, P(Ws), bp, P(acts), p(x), p(y), P(gW1), P(gb1) if bs else None, p(E1), p(ws), ws.numel(), st))
    gW2, gb2, E2 = _empty(wsh), _empty(bsh), _empty([(L,)])[0]
    _lib.call("pc_param_grads", d, Ws + bs + acts + [x, y], gW2 + gb2 + [E2, ws], st)
    _same(gW2 + gb2 + [E2], gW1 + gb1 + [E1])
    # train step
    A1, l1, E1, gW1, gb1 = _empty(ash), _empty([()])[0], _empty([(L,)])[0], _empty(wsh), _empty(bsh)
    _lib.check(L_.jpcb_pc_train_step(d, P(Ws), bp, p(x), p(y), P(A1), p(l1), p(E1), P(gW1), P(gb1) if bs else None, p(ws), ws.numel(), st))
    A2, l2, E2, gW2, gb2 = _empty(ash), _empty([()])[0], _empty([(L,)])[0], _empty(wsh), _empty(bsh)
    _lib.call("pc_train_step", d, Ws + bs + [x, y], A2 + [l2, E2] + gW2 + gb2 + [ws], st)
    _same(A2 + [l2, E2] + gW2 + gb2, A1 + [l1, E1] + gW1 + gb1)
    # hybrid PC: layer-wise regression of an amortiser with this descriptor
    D = list(net.dims)
    U = [torch.randn(B, D[l], device="cuda") for l in range(L)]
    T = [torch.randn(B, D[l + 1], device="cuda") for l in range(L)]
    gW1, gb1, E1 = _empty(wsh), _empty(bsh), _empty([(L,)])[0]
    _lib.check(L_.jpcb_hpc_param_grads(d, P(Ws), bp, P(U), P(T), P(gW1), P(gb1) if bs else None, p(E1), p(ws), ws.numel(), st))
    gW2, gb2, E2 = _empty(wsh), _empty(bsh), _empty([(L,)])[0]
    _lib.call("hpc_param_grads", d, Ws + bs + U + T, gW2 + gb2 + [E2, ws], st)
    _same(gW2 + gb2 + [E2], gW1 + gb1 + [E1])
    # ePC
    errs = [0.1 * torch.randn_like(a) for a in acts]
    F1, gE1, gW1, gb1 = _empty([()])[0], _empty(ash), _empty(wsh), _empty(bsh)
    _lib.check(L_.jpcb_epc_grads(d, P(Ws), bp, P(errs), p(x), p(y), p(F1), P(gE1), P(gW1), P(gb1) if bs else None, p(ws), ws.numel(), st))
    F2, gE2, gW2, gb2 = _empty([()])[0], _empty(ash), _empty(wsh), _empty(bsh)
    _lib.call("epc_grads", d, Ws + bs + errs + [x, y], [F2] + gE2 + gW2 + gb2 + [ws], st)
    _same([F2] + gE2 + gW2 + gb2, [F1] + gE1 + gW1 + gb1)
    torch.cuda.synchronize()


def test_pc_free_input_positional():
    """x = None: the state list has L + 1 entries and there is no x operand."""
    from jpc_b200 import _lib
    L_ = _lib.lib
    P, p = _lib.ptr_array, _lib.ptr
    net, acts = _pc(free=True)
    d, L = net.desc, net.L
    Ws, bs = _params(net)
    ws = _ws(L_.jpcb_pc_workspace_bytes(d))
    st = torch.cuda.current_stream().cuda_stream
    ash = [a.shape for a in acts]
    assert len(acts) == L + 1
    assert _lib.call_counts("pc_infer", d) == (L + len(bs) + (L + 1) + 1, (L + 1) + 1)
    F1, g1 = _empty([()])[0], _empty(ash)
    _lib.check(L_.jpcb_pc_activity_grad(d, P(Ws), P(bs), P(acts), None, p(net.y), p(F1), P(g1), p(ws), ws.numel(), st))
    F2, g2 = _empty([()])[0], _empty(ash)
    _lib.call("pc_activity_grad", d, Ws + bs + acts + [net.y], [F2] + g2 + [ws], st)
    _same([F2] + g2, [F1] + g1)
    S1 = [a.clone() for a in acts]
    _lib.check(L_.jpcb_pc_infer(d, P(Ws), P(bs), P(S1), None, p(net.y), None, p(ws), ws.numel(), st))
    S2 = _empty(ash)
    _lib.call("pc_infer", d, Ws + bs + acts + [net.y], S2 + [ws], st)
    _same(S2, S1)
    with pytest.raises(ValueError):   # an x operand too many
        _lib.call("pc_infer", d, Ws + bs + acts + [net.y, net.y], S2 + [ws], st)
    torch.cuda.synchronize()


@pytest.mark.parametrize("free", [False, True])
def test_bpc_entries_positional_equals_structured(free):
    import jpc_b200 as J
    from jpc_b200 import _bpc, _lib
    L_ = _lib.lib
    P, p = _lib.ptr_array, _lib.ptr
    J.set_compute_dtype("f32")
    d_x = 12 if free else 6
    td = O.make_mlp(0, d_x, 12, 3, 9, "tanh", use_bias=True)
    bu = O.make_mlp(1, 9, 12, 3, d_x, "tanh", use_bias=True)[::-1]
    x, y = data(B, d_x, 9, onehot_out=False)
    acts = devs(perturb(O.init_activities_with_ffwd(td, x)))
    net = _bpc._BpcNet(to_gpu_model(td), to_gpu_model(bu), None if free else dev(x), dev(y), skip_model=None, param_type="sp",
                       backward_energy_weight=0.7, forward_energy_weight=1.0, n_steps=3, dt=0.4)
    d, n = net.desc, net.H + 1
    Ws, Vs, bWs, bVs = (list(t) for t in net._keep)
    ws = _ws(L_.jpcb_bpc_workspace_bytes(d))
    st = net.stream()
    xs = [] if free else [net.x]
    ops = Ws + bWs + Vs + bVs + acts + xs + [net.y]
    assert _lib.call_counts("bpc_train_step", d) == (len(ops), n + 4 * n + 1 + 1)
    ash = [a.shape for a in acts]
    wsh, vsh, bwsh, bvsh = ([t.shape for t in q] for q in (Ws, Vs, bWs, bVs))
    head = (d, P(Ws), P(bWs), P(Vs), P(bVs))
    # energy pairs
    E1, E2 = _empty([(2 * n,)])[0], _empty([(2 * n,)])[0]
    _lib.check(L_.jpcb_bpc_energy(*head, P(acts), p(net.x), p(net.y), p(E1), p(ws), ws.numel(), st))
    _lib.call("bpc_energy", d, ops, [E2, ws], st)
    _same([E2], [E1])
    # activity gradient
    F1, g1 = _empty([()])[0], _empty(ash)
    _lib.check(L_.jpcb_bpc_activity_grad(*head, P(acts), p(net.x), p(net.y), p(F1), P(g1), p(ws), ws.numel(), st))
    F2, g2 = _empty([()])[0], _empty(ash)
    _lib.call("bpc_activity_grad", d, ops, [F2] + g2 + [ws], st)
    _same([F2] + g2[:-1], [F1] + g1[:-1])   # (the last slot is the unused dummy: no gradient is written)
    # inference
    S1 = [a.clone() for a in acts]
    _lib.check(L_.jpcb_bpc_infer(*head, P(S1), p(net.x), p(net.y), p(ws), ws.numel(), st))
    S2 = _empty(ash)
    _lib.call("bpc_infer", d, ops, S2 + [ws], st)
    _same(S2[:-1], S1[:-1])
    assert _rel(S1[0], acts[0]) > 1e-4
    # parameter gradients
    o1 = [_empty(wsh), _empty(bwsh), _empty(vsh), _empty(bvsh)]
    _lib.check(L_.jpcb_bpc_param_grads(*head, P(acts), p(net.x), p(net.y), *(P(q) for q in o1), p(E1), p(ws), ws.numel(), st))
    o2 = [_empty(wsh), _empty(bwsh), _empty(vsh), _empty(bvsh)]
    _lib.call("bpc_param_grads", d, ops, sum(o2, []) + [E2, ws], st)
    _same(sum(o2, []) + [E2], sum(o1, []) + [E1])
    # train step
    S1 = [a.clone() for a in acts]
    _lib.check(L_.jpcb_bpc_train_step(*head, P(S1), p(net.x), p(net.y), *(P(q) for q in o1), p(E1), p(ws), ws.numel(), st))
    S2, o2 = _empty(ash), [_empty(wsh), _empty(bwsh), _empty(vsh), _empty(bvsh)]
    _lib.call("bpc_train_step", d, ops, S2 + sum(o2, []) + [E2, ws], st)
    _same(S2[:-1] + sum(o2, []) + [E2], S1[:-1] + sum(o1, []) + [E1])
    torch.cuda.synchronize()


def test_positional_call_refuses_wrong_lists():
    from jpc_b200 import _lib
    net, acts = _pc()
    d = net.desc
    Ws, bs = _params(net)
    ws = _ws(1 << 20)
    st = torch.cuda.current_stream().cuda_stream
    with pytest.raises(ValueError):   # a result missing
        _lib.call("pc_energy", d, Ws + bs + acts + [net.x, net.y], [ws], st)
    with pytest.raises(ValueError):   # a PC descriptor handed to a bPC entry
        _lib.call("bpc_energy", d, Ws + bs + acts + [net.x, net.y], [acts[0], ws], st)
    with pytest.raises(ValueError):   # unknown entry id
        _lib.check(_lib.lib.jpcb_call(99, C.cast(C.pointer(d), C.c_void_p), C.sizeof(d), _lib.ptr_array(Ws), len(Ws),
                                      _lib.ptr_array([ws]), 1, ws.numel(), st))
