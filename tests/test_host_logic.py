"""Host-side bookkeeping that runs every train step (no GPU needed): flat buffers carved into views, the make_mlp fast
path of the model rebuild, pointer tables."""
import torch

from jpc_b200 import _flat, _lib, nn


def test_flat_layout_views_are_contiguous_disjoint_and_aligned():
    shapes = []
    for s in [(128, 784)] + [(128, 128)] * 5 + [(10, 128)]:
        shapes += [s, (s[0],)]          # interleaved weights / biases: runs of equal shapes at a constant spacing
    shapes += [(7,), ()]
    lay = _flat.layout(shapes)
    assert _flat.layout(shapes) is lay   # cached per shape list
    flat = torch.zeros(lay.total)
    views = lay.carve(flat)
    assert len(lay.runs) < len(shapes)   # (fewer as_strided calls than tensors)
    for i, (v, s) in enumerate(zip(views, shapes)):
        assert tuple(v.shape) == tuple(s) and v.is_contiguous()
        assert v.data_ptr() == flat.data_ptr() + 4 * lay.offs[i] and lay.offs[i] % _flat.ALIGN == 0
        v.add_(1.0)
    assert float(flat.sum()) == sum(lay.sizes)   # every element of every view touched exactly once: no overlap
    lead = _flat.layout([(4, 8)] * 3 + [(4, 2)]).carve(torch.zeros(4096), lead=(1,))
    assert [tuple(v.shape) for v in lead] == [(1, 4, 8)] * 3 + [(1, 4, 2)] and all(v.is_contiguous() for v in lead)
    outs = _flat.empty_like_list([torch.zeros(3, 5), torch.zeros(5), torch.zeros(3, 5)])
    assert [tuple(o.shape) for o in outs] == [(3, 5), (5,), (3, 5)]


def test_model_rebuild_fast_path_equals_the_generic_walk():
    for bias in (False, True):
        model = nn.make_mlp(0, 6, 8, 4, 3, "tanh", use_bias=bias, device="cpu")
        skips = nn.make_skip_model(4)
        leaves = [torch.full_like(p, float(i)) for i, p in enumerate(nn.tree_leaves((model, skips)))]
        fast_m, fast_s = nn.tree_unflatten_like((model, skips), leaves)
        assert type(fast_m) is nn.ModelList and isinstance(fast_m, list) and fast_m._jpcb is None
        again, _ = nn.tree_unflatten_like((fast_m, skips), leaves)        # a returned model goes back in: still the fast path
        assert type(again) is nn.ModelList and again[0].layers[1].weight is leaves[0]
        slow_m, slow_s = nn._unflatten((model, skips), iter(leaves))
        assert len(fast_m) == len(slow_m) == 4 and fast_s[1] is skips[1] and fast_s[0] is None
        for a, b in zip(nn.tree_leaves(fast_m), nn.tree_leaves(slow_m)):
            assert a is b
        for f, s, o in zip(fast_m, slow_m, model):
            assert type(f) is nn.Sequential and f.layers[0] is o.layers[0]           # activation objects are shared
            assert f.layers[1].in_features == o.layers[1].in_features and f.layers[1] is not o.layers[1]
            assert (f.layers[1].bias is None) == (not bias)
        assert nn.tree_unflatten_like(model, leaves)[2].layers[1].weight is leaves[2 * (2 if bias else 1)]
    # anything that is not the make_mlp form takes the generic walk
    odd = [nn.Linear(3, 2, key=0, device="cpu"), {"a": torch.zeros(2)}]
    new = nn.tree_unflatten_like(odd, [torch.ones(2, 3), torch.ones(2), torch.ones(2)])
    assert float(new[0].weight.sum()) == 6 and float(new[1]["a"].sum()) == 2


def test_pointer_table():
    ts = [torch.zeros(3), None, torch.zeros(2)]
    arr = _lib.ptr_array(ts)
    assert arr[0] == ts[0].data_ptr() and arr[1] is None and arr[2] == ts[2].data_ptr()
    assert len(_lib.ptr_array([])) == 1


def test_gradient_arena_layout_buckets_and_lazy_views():
    """The data-parallel exchange's flat arena (jpc_b200/_train.py:_Arena): segments aligned and disjoint, [E_layers, loss]
    at the tail, layer buckets contiguous and covering everything, per-parameter views only made on demand."""
    from jpc_b200._train import _Arena
    for bias in (False, True):
        ws = [(20, 12)] + [(20, 20)] * 5 + [(7, 20)]
        bs = [(s[0],) for s in ws] if bias else None
        arena = _Arena(weight_shapes=ws, bias_shapes=bs, device="cpu")
        assert "gW" not in arena.__dict__ and tuple(arena.E.shape) == (7,) and arena.loss.ndim == 0
        arena.flat.zero_()
        for n in (1, 2, 4, 7, 50):
            bk = arena.buckets(n)
            assert bk[0][0] == 0 and bk[0][2] == 0 and bk[-1][1] == 7 and bk[-1][3] == arena.flat.numel()
            for (l0, l1, f0, f1), (m0, m1, g0, g1) in zip(bk, bk[1:]):
                assert l1 == m0 and f1 == g0 and l0 < l1 and f0 < f1       # contiguous in layers and in the buffer
            assert len(bk) == min(n, 7)
        gW, gb = arena.gW, arena.gb                                          # carved now
        assert len(gW) == 7 and (gb is None) == (not bias)
        for l, g in enumerate(gW):
            assert tuple(g.shape) == ws[l] and g.data_ptr() == arena.flat.data_ptr() + 4 * arena.offs[arena.per_layer * l]
            g.add_(1.0)
        if bias:
            for g in gb:
                g.add_(1.0)
        arena.E.add_(1.0)
        arena.loss.add_(1.0)
        n_el = sum(a * b for a, b in ws) + (sum(s[0] for s in bs) if bias else 0) + 7 + 1
        assert float(arena.flat.sum()) == n_el                                # every element of every view exactly once
        pW, pb = arena.grad_ptr_arrays()
        assert pW[3] == gW[3].data_ptr() and ((pb is None) if not bias else pb[3] == gb[3].data_ptr())
