"""Evaluation helpers of jpc/_test.py: `test_discriminative_pc` (:24-73, feed-forward pass, loss and accuracy of
the output predictions) and `test_generative_pc` (:76-177, accuracy of the INPUT inferred by an unsupervised
solve from random-normal activities, plus the feed-forward output predictions).  Both are thin callers of the
ffwd-init and inference entry points."""
from __future__ import annotations

from ._core import init_activities_from_normal, init_activities_with_ffwd, solve_inference
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


def test_generative_pc(model, output, input, key, layer_sizes, batch_size, *, skip_model=None, loss_id="mse", param_type="sp",
                       sigma=0.05, ode_solver=None, max_t1=500, dt=None, stepsize_controller=None, weight_decay=0.0,
                       spectral_penalty=0.0, activity_decay=0.0):
    """jpc/_test.py:76-177: (accuracy of the inferred input, feed-forward output predictions)."""
    _check_param_type(param_type)
    activities = init_activities_from_normal(key=key, layer_sizes=layer_sizes, mode="unsupervised", batch_size=batch_size,
                                             sigma=sigma, device=output.device)
    input_preds = solve_inference((model, skip_model), activities, output, loss_id=loss_id, param_type=param_type,
                                  solver=ode_solver, max_t1=max_t1, dt=dt, stepsize_controller=stepsize_controller,
                                  weight_decay=weight_decay, spectral_penalty=spectral_penalty,
                                  activity_decay=activity_decay)[0][0]
    input_acc = compute_accuracy(input, input_preds)
    output_preds = init_activities_with_ffwd(model=model, input=input, skip_model=skip_model, param_type=param_type)[-1]
    return input_acc, output_preds


test_generative_pc.__test__ = False
