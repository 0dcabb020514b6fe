"""diffrax-shaped stand-ins for the solver / step-size-controller objects the reference's
`solve_inference` and `make_pc_step` take (jpc/_core/_infer.py:7-17,28-33;
experiments/library_paper/train_mlp.py:81-84).  They carry configuration only; the stepping
itself happens inside the CUDA library (jpcb_pc_infer)."""
from __future__ import annotations


class Euler:
    """diffrax.Euler: y1 = y0 + dt * f(t0, y0)."""

    def __repr__(self):
        return "Euler()"


class Heun:
    """diffrax.Heun: explicit trapezoidal rule, tableau c=[1], a21=1, b=[1/2, 1/2]."""

    def __repr__(self):
        return "Heun()"


class ConstantStepSize:
    """diffrax.ConstantStepSize: every step is dt0, the last one clipped to t1."""

    def __repr__(self):
        return "ConstantStepSize()"


class PIDController:
    """diffrax.PIDController(rtol, atol): accepted for signature compatibility; adaptive stepping
    is a 'next' row of the scope table and raises NotImplementedError in solve_inference."""

    def __init__(self, rtol: float = 1e-3, atol: float = 1e-3, **kw):
        self.rtol, self.atol = rtol, atol
        self.kw = kw

    def __repr__(self):
        return f"PIDController(rtol={self.rtol}, atol={self.atol})"
