"""Flat fp32 device buffers carved into per-tensor views -- host-side bookkeeping only.

The functional, reference-shaped API returns fresh tensors every step (new parameters, gradients, activities).  For a
128-layer network that used to be several hundred `torch.empty` / `as_strided` calls per step -- more host time than the
GPU needs for the step.  Here every such family of outputs is ONE allocation, and the views are
made with one `as_strided(...).unbind(0)` per run of equally shaped, equally spaced tensors (a 128-layer residual MLP has
3 such runs).  One allocation per family also makes the device addresses repeat from step to step (the caching allocator
hands the same blocks back), which is what lets the library replay its launch plans (csrc/plan.cu).
"""
from __future__ import annotations

import torch

ALIGN = 64  # segments start on 256-byte boundaries (the kernels use 16-byte vector accesses)

_LAYOUTS: dict = {}


def _contig(shape):
    st, acc = [], 1
    for n in reversed(shape):
        st.append(acc)
        acc *= n
    return tuple(reversed(st))


class Layout:
    """Offsets (in elements) of tensors of the given shapes inside one flat buffer + the runs to carve them by."""

    __slots__ = ("shapes", "offs", "sizes", "total", "runs", "_numel_table")

    def __init__(self, shapes):
        self.shapes = tuple(tuple(int(v) for v in s) for s in shapes)
        self.sizes, self.offs, off = [], [], 0
        for s in self.shapes:
            n = 1
            for v in s:
                n *= v
            self.sizes.append(n)
            self.offs.append(off)
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        self.total = max(off, 1)
        self._numel_table = None  # (ctypes table of `sizes`, made on first use by the fused parameter update)
        # runs: (indices, shape, first offset, spacing): equal shapes at a constant spacing, in index order
        by_shape: dict = {}
        for i, s in enumerate(self.shapes):
            by_shape.setdefault(s, []).append(i)
        self.runs = []
        for s, idx in by_shape.items():
            k = 0
            while k < len(idx):
                j = k + 1
                if j < len(idx):
                    step = self.offs[idx[j]] - self.offs[idx[k]]
                    while j < len(idx) and self.offs[idx[j]] - self.offs[idx[j - 1]] == step:
                        j += 1
                else:
                    step = 0
                self.runs.append((tuple(idx[k:j]), s, self.offs[idx[k]], step))
                k = j

    def carve(self, flat: torch.Tensor, lead: tuple = ()):
        """Views of `flat`, one per shape (each with `lead` singleton dims in front), in order."""
        out = [None] * len(self.shapes)
        for idx, s, off, step in self.runs:
            full = tuple(lead) + s
            st = _contig(full)
            if len(idx) == 1:
                out[idx[0]] = flat.as_strided(full, st, off)
            else:
                for i, v in zip(idx, flat.as_strided((len(idx),) + full, (step,) + st, off).unbind(0)):
                    out[i] = v
        return out


def layout(shapes) -> Layout:
    key = tuple(shapes)
    lay = _LAYOUTS.get(key)
    if lay is None:
        if len(_LAYOUTS) > 512:
            _LAYOUTS.clear()
        lay = _LAYOUTS[key] = Layout(key)
    return lay


def empty_like_list(tensors):
    """Fresh tensors with the shapes of `tensors` (contiguous fp32 on one device), carved from ONE allocation."""
    lay = layout([t.shape for t in tensors])
    flat = torch.empty(lay.total, dtype=torch.float32, device=tensors[0].device)
    return lay.carve(flat)
