"""Evaluation helper of jpc/_test.py:24-73 (`test_discriminative_pc`): feed-forward pass (the ffwd-init
kernel), loss and accuracy of the output predictions."""
from __future__ import annotations

from ._core import init_activities_with_ffwd
from ._errors import _check_param_type
from ._utils import compute_accuracy, cross_entropy_loss, mse_loss


def test_discriminative_pc(model, output, input, *, skip_model=None, loss="mse", param_type="sp"):
    """jpc/_test.py:24-73: (test loss, accuracy) of the feed-forward predictions."""
    _check_param_type(param_type)
    preds = init_activities_with_ffwd(model=model, input=input, skip_model=skip_model, param_type=param_type)[-1]
    if loss == "mse":
        loss = mse_loss(preds, output)
    elif loss == "ce":
        loss = cross_entropy_loss(preds, output)
    acc = compute_accuracy(output, preds)
    return loss, acc


test_discriminative_pc.__test__ = False  # (not a pytest test, despite the reference's name)
