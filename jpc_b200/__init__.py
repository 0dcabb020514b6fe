"""jpc_b200 -- B200-native (sm_100a CUDA) drop-in for the predictive-coding train-step path of
thebuckleylab/jpc.  Same function names as `jpc` for the path it covers (jpc/__init__.py:3-61);
everything is computed by libjpc_b200.so through the C-ABI in include/jpc_b200.h.  Importing this
package loads (and if needed builds) that library and raises if it cannot: there is no fallback.
"""
from . import _lib  # noqa: F401  (loads libjpc_b200.so or raises)
from ._lib import plan_cache_clear, plan_cache_configure, plan_cache_stats, set_reproducible
from ._core import (
    _get_param_scalings,
    compute_pc_activity_grad,
    compute_pc_param_grads,
    get_compute_dtype,
    init_activities_from_normal,
    init_activities_with_ffwd,
    neg_pc_activity_grad,
    pc_energy_fn,
    set_compute_dtype,
    solve_inference,
    update_pc_activities,
    update_pc_params,
)
from ._bpc import (bpc_energy_fn, bpc_train_step, compute_bpc_activity_grad, compute_bpc_param_grads, solve_bpc_inference,
                   update_bpc_activities, update_bpc_params)
from ._errors import _check_param_type
from ._eval import test_discriminative_pc, test_generative_pc
from ._epc import (compute_epc_error_grad, compute_epc_param_grads, epc_energy_fn, init_epc_errors, update_epc_errors,
                   update_epc_params)
from ._hpc import (compute_hpc_param_grads, hpc_energy_fn, init_activities_with_amort, make_hpc_step, test_hpc)
from ._train import make_pc_step, set_dp_buckets, set_grad_allreduce_dtype
from ._graph import PCTrainer
from ._utils import (compute_accuracy, compute_activity_norms, compute_infer_energies, compute_param_norms,
                     cross_entropy_loss, get_t_max, mse_loss)
from .nn import get_act_fn, make_mlp, make_skip_model
from .solvers import ConstantStepSize, Euler, Heun, PIDController
from . import nn, optim  # noqa: F401

# legacy spellings still used by the reference's README and experiment scripts
# (README.md:126,138; experiments/library_paper/train_mlp.py:130,142)
update_activities = update_pc_activities
update_params = update_pc_params

__version__ = "0.2.0"
