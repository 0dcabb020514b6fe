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
    """diffrax.PIDController(rtol, atol, pcoeff=0, icoeff=1, dcoeff=0, safety=0.9, factormin=0.2, factormax=10):
    adaptive step sizes from the solver's embedded error estimate.  The trial steps run in the CUDA library
    (jpcb_pc_heun_trial); this object holds the controller's scalar arithmetic, restated from upstream diffrax
    0.6-0.7 (third-party, not under /root/reference; SURVEY Appendix B -- "recalled", unverifiable offline):

        scaled_error = rms_norm(y_err / (atol + rtol * max(|y0|, |y1|)));   accept iff scaled_error < 1
        factor = clip(safety * scaled_error**(-icoeff/order) * prev**(..pcoeff, dcoeff terms..),
                      min = 1 if accepted else factormin, max = factormax);   dt_next = dt * factor

    `order` is the solver's error order (2 for Heun)."""

    def __init__(self, rtol: float = 1e-3, atol: float = 1e-3, pcoeff: float = 0.0, icoeff: float = 1.0, dcoeff: float = 0.0,
                 safety: float = 0.9, factormin: float = 0.2, factormax: float = 10.0, **kw):
        self.rtol, self.atol = float(rtol), float(atol)
        self.pcoeff, self.icoeff, self.dcoeff = float(pcoeff), float(icoeff), float(dcoeff)
        self.safety, self.factormin, self.factormax = float(safety), float(factormin), float(factormax)
        self.kw = kw

    def init_state(self):
        return (1.0, 1.0)  # inverse scaled errors of the two previous steps

    def adapt(self, dt: float, scaled_error: float, order: float, state):
        """-> (accepted, next dt, next state)"""
        keep = scaled_error < 1.0
        inv = float("inf") if scaled_error == 0.0 else 1.0 / scaled_error
        prev, prev_prev = state
        c1 = (self.icoeff + self.pcoeff + self.dcoeff) / order
        c2 = -(self.pcoeff + 2.0 * self.dcoeff) / order
        c3 = self.dcoeff / order
        f1 = 1.0 if c1 == 0.0 else inv ** c1
        f2 = 1.0 if c2 == 0.0 else prev ** c2
        f3 = 1.0 if c3 == 0.0 else prev_prev ** c3
        fmin = 1.0 if keep else self.factormin
        factor = min(max(self.safety * f1 * f2 * f3, fmin), self.factormax)
        new_state = (inv, prev) if keep else state
        return keep, dt * factor, new_state

    def __repr__(self):
        return f"PIDController(rtol={self.rtol}, atol={self.atol})"
