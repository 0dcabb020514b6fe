/* jpc_b200.h -- C-ABI of the B200-native predictive-coding (PC) train-step path.
 *
 * This is the drop-in boundary for the hot path of thebuckleylab/jpc (reference citations are
 * relative to the reference repository root).  The reference has no FFI of its own (it is pure
 * Python/JAX); each entry point below replaces the XLA program that one reference function
 * lowers to, and is what a `jax.ffi` custom call (or the ctypes binding in jpc_b200/_lib.py)
 * binds.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *  - Plain pointers and sizes only; no torch / XLA types.  Every data pointer is a DEVICE pointer
 *    unless stated otherwise; arrays of pointers (`W`, `b`, `A`, ...) are HOST arrays of device
 *    pointers.  All tensors are dense row-major fp32: activities [batch_local, d], weights
 *    [d_out, d_in] (equinox nn.Linear convention), biases [d_out].
 *  - The caller owns every buffer, including the workspace (`jpcb_pc_workspace_bytes`).  The
 *    library allocates nothing, never synchronises the device, and enqueues only on `stream`.
 *  - Return value: 0 on success, a JPCB_E* code otherwise; `jpcb_last_error()` gives the message
 *    (thread-local).  There is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with JPCB_ECUDA.
 *  - Data parallelism: each rank passes its `batch_local` rows and the GLOBAL batch size used for
 *    the 1/B normalisation (jpc/_core/_energies.py:85,153); energies, losses and parameter
 *    gradients returned are this rank's partial sums, to be all-reduced with SUM by the caller.
 */
#ifndef JPC_B200_H_
#define JPC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JPCB_VERSION 102
#define JPCB_MAX_LAYERS 256

/* error codes */
#define JPCB_OK 0
#define JPCB_EINVAL 1   /* bad descriptor / argument (ValueError in the reference) */
#define JPCB_ENOTIMPL 2 /* valid in the reference but not on this path (NotImplementedError) */
#define JPCB_EWORKSPACE 3
#define JPCB_ECUDA 4

/* GEMM operand arithmetic.  F32: fp32 CUDA-core FMA (1e-5 parity configs).  BF16: tcgen05 tensor
 * cores, bf16 operands, fp32 accumulation in TMEM, fp32 master activities (2e-2 parity). */
#define JPCB_DTYPE_F32 0
#define JPCB_DTYPE_BF16 1
/* fp32 parity on the tensor cores (bPC entry points): every GEMM operand is a 3-way bf16 split hi + mid + lo (24
 * mantissa bits) and every GEMM runs the six products hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi into one fp32
 * accumulator; activities, errors and epilogue arithmetic stay fp32 with accurate activations (1e-5 parity). */
#define JPCB_DTYPE_F32X3 2

/* activation ids, order of jpc/_utils.py:13-15 */
#define JPCB_ACT_LINEAR 0
#define JPCB_ACT_TANH 1
#define JPCB_ACT_HARD_TANH 2
#define JPCB_ACT_RELU 3
#define JPCB_ACT_LEAKY_RELU 4
#define JPCB_ACT_GELU 5
#define JPCB_ACT_SELU 6
#define JPCB_ACT_SILU 7

#define JPCB_LOSS_MSE 0
#define JPCB_LOSS_CE 1

#define JPCB_SOLVER_EULER 0 /* diffrax.Euler  + ConstantStepSize */
#define JPCB_SOLVER_HEUN 1  /* diffrax.Heun   + ConstantStepSize */

/* Descriptor of one PC network of dense layers, l = 0..n_layers-1:
 *   f_l(u) = post_act_l( W_l act_l(u) + b_l ).
 * The reference's make_mlp builds Form A, `Sequential([Lambda(act), Linear])`: post_act = identity
 * (jpc/_utils.py:87-106); its hand-built example networks are Form B, `Sequential([Linear, Lambda(act)])`: act = identity
 * (examples/jpc_from_scratch.ipynb cell 12, examples/discriminative_pc.ipynb cell 8). */
typedef struct jpcb_pc_desc {
  int32_t n_layers;                  /* L = len(model) >= 2 */
  int32_t batch_local;               /* rows held by this rank */
  int32_t batch_global;              /* B of the 1/B energy normalisation */
  int32_t dims[JPCB_MAX_LAYERS + 1]; /* d_0 .. d_L */
  int32_t act[JPCB_MAX_LAYERS];      /* activation applied to the INPUT of layer l */
  int32_t has_bias;                  /* all layers carry a bias (make_mlp use_bias) */
  uint8_t skip[JPCB_MAX_LAYERS];     /* identity skip at layer l (jpc/_utils.py:122-126) */
  float scalings[JPCB_MAX_LAYERS];   /* a_l of _get_param_scalings (jpc/_core/_energies.py:879-933) */
  int32_t dtype;                     /* JPCB_DTYPE_* */
  int32_t loss;                      /* JPCB_LOSS_*: output-layer energy (mse, or cross-entropy of softmax(logits)) */
  float output_energy_scaling;       /* lambda; 1 when the reference passes None */
  float activity_decay;
  int32_t solver;                    /* JPCB_SOLVER_* */
  int32_t n_steps;                   /* T fixed steps */
  float dt;                          /* step size of steps 0..T-2 */
  float dt_last;                     /* step size of the last step (ConstantStepSize clips to t1) */
  int32_t free_input;                /* unsupervised PC (`x = None` in the reference, jpc/_core/_energies.py:86,
                                      * 123-124): there is no input; z_0 is a free activity.  Every activity list
                                      * (A, gA, A_new) then has n_layers + 1 entries [z_0, z_1 .. z_{L-1}, dummy]
                                      * exactly like the reference's, and the `x` argument must be NULL.  scalings[0]
                                      * is applied as given: the reference's PC branch is unscaled and only runs with
                                      * "sp" (callers pass 1); its ePC branch scales the first layer like any other.
                                      * jpcb_pc_ffwd_init / jpcb_pc_train_step need an input. */
  int32_t post_act[JPCB_MAX_LAYERS]; /* activation applied to the OUTPUT of layer l's linear map (Form B); 0 = none.
                                      * With pre = W_l act_l(u_l) + b_l:  e_{l+1} = t - a_l post(pre) - s_l u_l, and the
                                      * error that travels through W_l (activity and weight gradients) is post'(pre) . e_{l+1}
                                      * (SURVEY A.2 "Form B").  Supervised PC entry points only (not free_input, bPC, hybrid
                                      * PC, ePC); a cross-entropy output layer must have post_act = 0. */
} jpcb_pc_desc;

int jpcb_version(void);
const char* jpcb_last_error(void);
/* Number of CUDA kernels this library has enqueued in this process (bench.py's gpu_launches). */
unsigned long long jpcb_launch_count(void);

/* Energies and losses of the PC entry points (mse) by FIXED-ORDER sums over the stored errors instead of the fused
 * epilogues' floating-point atomics: the values then depend on the data only (bit-for-bit reproducible run to run; the
 * default accumulation differs at 1e-7 with the order blocks happen to finish in).  Costs two small launches per energy
 * evaluation; on the plain bf16 path the sums read the bf16 error copies.  Cross-entropy energies, the adaptive
 * controller's norms and the bPC energies keep the atomics.  Process-wide; default off. */
int jpcb_set_reproducible(int on);

/* Bytes of caller-provided device workspace any entry point below may use for `desc`. */
size_t jpcb_pc_workspace_bytes(const jpcb_pc_desc* desc);

/* init_activities_with_ffwd  (jpc/_core/_init.py:12-82).
 * A_out[l] (l = 0..L-1) receives z_{l+1} = a_l f_l(z_l) (+ z_l where skip[l]); A_out[L-1] is the
 * network output.  If `loss_out` and `y` are non-NULL, *loss_out (device float) receives this
 * rank's part of mse_loss(A_out[L-1], y) = 1/(2B) sum (y - out)^2, or of cross_entropy_loss when
 * desc->loss is JPCB_LOSS_CE  (jpc/_utils.py:129-136, jpc/_train.py:159-164). */
int jpcb_pc_ffwd_init(const jpcb_pc_desc* desc, const float* const* W, const float* const* b,
                      const float* x, float* const* A_out, const float* y, float* loss_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* pc_energy_fn(..., record_layers=True)  (jpc/_core/_energies.py:13-158).
 * E_layers: device float[L], order [output layer, hidden net_l = 1..L-2, first layer], each
 * divided by batch_global (regularisers excluded, as in the reference's record_layers branch). */
int jpcb_pc_energy(const jpcb_pc_desc* desc, const float* const* W, const float* const* b,
                   const float* const* A, const float* x, const float* y, float* E_layers,
                   void* workspace, size_t workspace_bytes, void* stream);

/* compute_pc_activity_grad / neg_pc_activity_grad  (jpc/_core/_grads.py:22-167).
 * gA[l] receives dF/dA[l] for l = 0..L-1 (gA[L-1], the dummy slot, is activity_decay*A/B).
 * F (device float, may be NULL) receives the total energy incl. the activity-decay term. */
int jpcb_pc_activity_grad(const jpcb_pc_desc* desc, const float* const* W, const float* const* b,
                          const float* const* A, const float* x, const float* y, float* F,
                          float* const* gA, void* workspace, size_t workspace_bytes, void* stream);

/* solve_inference with a fixed-step solver  (jpc/_core/_infer.py:20-133): desc->n_steps steps of
 * Euler/Heun on dA/dt = -dF/dA, in place on A.  All L layers x T steps run without returning to
 * the host.  `sumsq_out` (device float[2], may be NULL) receives {sum A^2, numel} after the last
 * step so the caller can evaluate the reference's steady-state Event (jpc/_core/_infer.py:136-145). */
int jpcb_pc_infer(const jpcb_pc_desc* desc, const float* const* W, const float* const* b,
                  float* const* A, const float* x, const float* y, float* sumsq_out,
                  void* workspace, size_t workspace_bytes, void* stream);

/* One TRIAL step of diffrax.Heun of size dt from state A (which is not modified), for solve_inference with an
 * adaptive step-size controller (jpc/_core/_infer.py:109-132 with PIDController; the accept / reject / next-dt
 * logic is scalar host code in the caller, as it is host-side controller code in diffrax).  desc->solver must be
 * JPCB_SOLVER_HEUN.  A_new[l] (l = 0..L-1) receives the candidate state; out2 (device float[2]) receives
 *   out2[0] = sum over every activity entry of ( y_err / (atol + rtol * max(|y0|, |y1|)) )^2,
 *             y_err = dt (k1 - k2) / 2 the embedded error estimate (the dummy slot contributes 0),
 *   out2[1] = sum of A_new^2 over all L slots (the reference's steady-state Event, _infer.py:136-145). */
int jpcb_pc_heun_trial(const jpcb_pc_desc* desc, const float* const* W, const float* const* b,
                       const float* const* A, const float* x, const float* y, float dt, float rtol,
                       float atol, float* const* A_new, float* out2, void* workspace,
                       size_t workspace_bytes, void* stream);

/* compute_pc_param_grads  (jpc/_core/_grads.py:436-499).  gW[l]: [d_{l+1}, d_l]; gb[l]:
 * [d_{l+1}] (ignored when !has_bias; gb may then be NULL).  Partial sums over local rows,
 * already divided by batch_global.  E_layers (may be NULL) as in jpcb_pc_energy, for free. */
int jpcb_pc_param_grads(const jpcb_pc_desc* desc, const float* const* W, const float* const* b,
                        const float* const* A, const float* x, const float* y, float* const* gW,
                        float* const* gb, float* E_layers, void* workspace, size_t workspace_bytes,
                        void* stream);

/* make_pc_step up to (excluding) the optimiser  (jpc/_train.py:135-232): ffwd init, loss at the
 * init, T-step inference, per-layer energies at the solution, parameter gradients.
 * gW == NULL: everything but the parameter gradients; the prediction errors and operand copies at the
 * solution stay in `workspace` for jpcb_pc_param_grads_range (data-parallel callers that exchange the
 * gradients bucket by bucket while the remaining buckets are still being computed). */
int jpcb_pc_train_step(const jpcb_pc_desc* desc, const float* const* W, const float* const* b,
                       const float* x, const float* y, float* const* A_out, float* loss_out,
                       float* E_layers, float* const* gW, float* const* gb, void* workspace,
                       size_t workspace_bytes, void* stream);

/* compute_pc_param_grads (jpc/_core/_grads.py:487-499) for layers [layer_lo, layer_hi) only, from
 * the prediction errors / operand copies that the LAST jpcb_pc_train_step or jpcb_pc_param_grads call
 * with the same descriptor left in `workspace` (nothing else may have used the workspace in between).
 * gW / gb are the full per-layer pointer lists; only the entries of the range are written.  One
 * grouped launch per call: a caller records an event after each call and starts that bucket's
 * all-reduce while the next bucket's GEMMs run. */
int jpcb_pc_param_grads_range(const jpcb_pc_desc* desc, const float* const* W, const float* const* b,
                              const float* x, int layer_lo, int layer_hi, float* const* gW,
                              float* const* gb, void* workspace, size_t workspace_bytes, void* stream);

/* ---- bidirectional PC (jpc/_core/_energies.py:246-357) ---------------------------------------
 * H = td.n_layers - 1.  Top-down model W_j (j = 0..H): [td.dims[j+1], td.dims[j]], maps u_j (x for
 * j = 0, else z_{j-1}) to a prediction of t_j (z_j, or y for j = H).  Bottom-up model V_j (j = 0..H):
 * [td.dims[j], td.dims[j+1]], maps v_j (z_j, or y for j = H) to a prediction of tau_j (x for j = 0,
 * else z_{j-1}).  Activities A[0..H-1] = z_0..z_{H-1} with z_k of width td.dims[k+1]; A[H] is the
 * unused dummy slot ([batch, td.dims[H+1]]).  x: [batch, td.dims[0]], y: [batch, td.dims[H+1]].
 * td.skip[j] (1 <= j <= H-1) applies to both directions, td.scalings is ignored (the reference
 * validates param_type but applies no scalings, _energies.py:314).  td.free_input = 1 is the reference's
 * `x = None` (x_eff = activities[0], _energies.py:324): pass x = NULL, A[0] is both the input of W_0 and the
 * target of V_0 (needs td.dims[0] == td.dims[1]) and its gradient gains the two extra terms.  Euler
 * steps; td.dtype selects fp32 CUDA cores, bf16 tensor cores (needs td.act[k+1] == bu_act[k] and dims / batch
 * multiples of 8) or fp32 parity on tensor cores (JPCB_DTYPE_F32X3, same layout rules). */
typedef struct jpcb_bpc_desc {
  jpcb_pc_desc td;                 /* top-down model: dims d_0(x) .. d_{H+1}(y); solver fields used */
  int32_t bu_act[JPCB_MAX_LAYERS]; /* activation applied to the input of bottom_up_model[l] */
  int32_t bu_has_bias;
  float forward_energy_weight;     /* alpha_f */
  float backward_energy_weight;    /* alpha_b */
} jpcb_bpc_desc;

size_t jpcb_bpc_workspace_bytes(const jpcb_bpc_desc* desc);

/* bpc_energy_fn(record_layers=True): E_pairs = device float[2*(H+1)], reference order
 * [(e_out, delta_0), (e_0, delta'_0), ...] / batch  (jpc/_core/_energies.py:322-355). */
int jpcb_bpc_energy(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW,
                    const float* const* V, const float* const* bV, const float* const* A,
                    const float* x, const float* y, float* E_pairs, void* workspace,
                    size_t workspace_bytes, void* stream);

/* compute_bpc_activity_grad (jpc/_core/_grads.py:170-226) */
int jpcb_bpc_activity_grad(const jpcb_bpc_desc* desc, const float* const* W,
                           const float* const* bW, const float* const* V, const float* const* bV,
                           const float* const* A, const float* x, const float* y, float* F,
                           float* const* gA, void* workspace, size_t workspace_bytes, void* stream);

/* td.n_steps fused Euler steps (== update_bpc_activities with optax.sgd(dt), looped;
 * jpc/_core/_updates.py:206-281), in place on A. */
int jpcb_bpc_infer(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW,
                   const float* const* V, const float* const* bV, float* const* A, const float* x,
                   const float* y, void* workspace, size_t workspace_bytes, void* stream);

/* compute_bpc_param_grads (jpc/_core/_grads.py:544-612) */
int jpcb_bpc_param_grads(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW,
                         const float* const* V, const float* const* bV, const float* const* A,
                         const float* x, const float* y, float* const* gW, float* const* gbW,
                         float* const* gV, float* const* gbV, float* E_pairs, void* workspace,
                         size_t workspace_bytes, void* stream);

/* The bPC train step of the reference's example loop (examples/bidirectional_pc.ipynb cell 8) after the activity
 * initialisation: td.n_steps Euler steps in place on A (== jpcb_bpc_infer), then both models' parameter gradients and the
 * energy pairs at the solution (== jpcb_bpc_param_grads) -- as ONE call: the operands are prepared once and the operand
 * copies the last Euler step's epilogues wrote feed the weight-gradient GEMMs. */
int jpcb_bpc_train_step(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW,
                        const float* const* V, const float* const* bV, float* const* A, const float* x,
                        const float* y, float* const* gW, float* const* gbW, float* const* gV,
                        float* const* gbV, float* E_pairs, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ---- hybrid PC: the amortiser's layer-wise regression -------------------------------
 * hpc_energy_fn (jpc/_core/_energies.py:161-243) and compute_hpc_param_grads (jpc/_core/_grads.py:502-541):
 * for every layer l of the amortiser `desc` describes (unit scalings, no skips, mse),
 *   e_l = T[l] - (W_l act_l(U[l]) + b_l),   E_l = sum e_l^2 / B   (no 1/2 in this energy),
 *   gW[l] = -(2/B) e_l^T act_l(U[l]),       gb[l] = -(2/B) sum_rows e_l.
 * U[l] = the layer's input (U[0] = the amortiser's input, U[l] = its own feed-forward guess of layer l-1),
 * T[l] = the layer's target (the generator's equilibrated activity; T[L-1] = the amortiser's target).
 * E_layers (nullable): device float[L] in pc_energy_fn's order [last layer, 1..L-2, first layer].
 * gW / gb nullable as a pair (energy only). */
int jpcb_hpc_param_grads(const jpcb_pc_desc* desc, const float* const* W, const float* const* b,
                         const float* const* U, const float* const* T, float* const* gW,
                         float* const* gb, float* E_layers, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- error-reparameterised PC, ePC -------------------------------------------------
 * epc_energy_fn (jpc/_core/_energies.py:360-458), compute_epc_error_grad / compute_epc_param_grads
 * (jpc/_core/_grads.py:857-949) for supervised networks.  The state is the error list
 * E = [eps_1 .. eps_{L-1}, eps_L] (shapes of the PC activity list); activities follow from an error-perturbed
 * forward pass z_{l+1} = a_l f_l(z_l) + eps_{l+1} (+ z_l where skip[l]), z_L = a f(z_{L-1}) - eps_L, and
 *   F = ( sum_{l<L} 1/2 |eps_l|^2 + 1/2 |y - z_L|^2  [or the cross-entropy of z_L] ) / B.
 * One call = the forward chain (L dependent GEMMs), and if gradients are requested the backward chain
 * rho_l = act_l'(z_l) . (a_l rho_{l+1} W_l) + s_l rho_{l+1} from rho_L = y - z_L (L-1 dependent GEMMs):
 *   gE[l-1] = (eps_l - rho_l) / B  (l < L),   gE[L-1] = rho_L / B,   gW[l] = -(a_l/B) rho_{l+1}^T act_l(z_l).
 * F, gE, (gW, gb) are each nullable.  Workspace: jpcb_epc_workspace_bytes(desc) (larger than the PC one).
 * desc->free_input (x = None in the reference): E and gE have L + 1 entries, E[0] plays z_0, x must be NULL, and -- as in
 * the reference, whose penalty loop still covers the first L - 1 list entries -- z_0 is penalised and eps_{L-1} is not. */
size_t jpcb_epc_workspace_bytes(const jpcb_pc_desc* desc);
int jpcb_epc_grads(const jpcb_pc_desc* desc, const float* const* W, const float* const* b,
                   const float* const* E, const float* x, const float* y, float* F, float* const* gE,
                   float* const* gW, float* const* gb, void* workspace, size_t workspace_bytes,
                   void* stream);

/* ---- parameter update ----------------------------------------------------------
 * The last two lines of the reference's train step, `optim.update(updates=param_grads, state, params)`
 * followed by `eqx.apply_updates` (jpc/_train.py:234-242; jpc/_core/_updates.py:189-194), for the two
 * optimisers its examples use.  optax is a third-party dependency (pins: SURVEY 8c); restated:
 *   optax.sgd(lr):                p' = p - lr g
 *   optax.adam(lr, b1, b2, eps):  mu' = b1 mu + (1-b1) g;  nu' = b2 nu + (1-b2) g^2;
 *                                 p' = p - lr (mu'/(1-b1^t)) / (sqrt(nu'/(1-b2^t)) + eps),  t = step >= 1
 * One pass over `n_tensors` device tensors of `numel[i]` floats each (pointer tables are HOST arrays of
 * device pointers).  Hyper-parameters are doubles: 1-b, 1-b^t are formed in double and rounded once to fp32,
 * as optax does with Python floats.  Outputs may alias their inputs (in place) or be fresh buffers (the reference's
 * functional convention). */
int jpcb_optim_sgd(int n_tensors, const long long* numel, const float* const* params,
                   const float* const* grads, double learning_rate, float* const* params_out,
                   void* stream);
int jpcb_optim_adam(int n_tensors, const long long* numel, const float* const* params,
                    const float* const* grads, const float* const* mu, const float* const* nu,
                    double learning_rate, double b1, double b2, double eps, int step,
                    float* const* params_out, float* const* mu_out, float* const* nu_out, void* stream);

/* ---- weight regularisers of pc_energy_fn (jpc/_core/_energies.py:127-143) ------------------------------------------
 * The reference applies `weight_decay` and `spectral_penalty` to the layers that expose `.weight` directly -- bare
 * nn.Linear layers (the last layer of its hand-built example networks); for make_mlp models both are no-ops:
 *   F_reg       = weight_decay/2 sum_i |W_i|^2 + spectral_penalty/2 sum_i |W_i^T W_i - I|^2     (not divided by the batch)
 *   dF_reg/dW_i = weight_decay W_i + 2 spectral_penalty W_i (W_i^T W_i - I)
 * n tensors W[i]: [rows[i], cols[i]] (host arrays).  F_inout (device float, nullable) receives F_reg, or has it added when
 * accumulate_F != 0; gW_inout (nullable list; entries nullable) is ACCUMULATED into.  In data-parallel runs the caller adds
 * these terms once, after the gradient exchange. */
size_t jpcb_pc_weight_reg_workspace_bytes(int n, const int* rows, const int* cols);
int jpcb_pc_weight_reg(int n, const float* const* W, const int* rows, const int* cols, float weight_decay,
                       float spectral_penalty, float* F_inout, int accumulate_F, float* const* gW_inout,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- launch plans ("build once, replay") ---------------------------------------------
 * Every compute entry point above only ENQUEUES, and a call with identical arguments -- descriptor, every pointer
 * (including the contents of the pointer lists), workspace, stream, device -- enqueues an identical kernel sequence.
 * The reference pays the corresponding cost once, when `eqx.filter_jit` traces `make_pc_step` (jpc/_train.py:35); here
 * the library keeps a small cache of instantiated CUDA graphs: the `min_calls`-th time an identical call is seen its
 * sequence is captured on a library-owned stream and from then on the call is one cudaGraphLaunch on the caller's
 * stream (the functional host API allocates its outputs from a caching allocator, so its pointer sets repeat with a
 * short period).  Results are identical; jpcb_launch_count() keeps counting the kernels inside replayed graphs.
 * Calls made while the caller is itself capturing `stream` bypass the cache.  Planned entry points: jpcb_pc_ffwd_init,
 * jpcb_pc_energy, jpcb_pc_activity_grad, jpcb_pc_infer (sumsq_out == NULL), jpcb_pc_param_grads, jpcb_pc_train_step,
 * jpcb_bpc_energy, jpcb_bpc_activity_grad, jpcb_bpc_infer, jpcb_bpc_param_grads, jpcb_bpc_train_step.
 * Defaults: 16 plans, min_calls 2 (environment: JPCB_PLAN=0 disables, JPCB_PLAN=n plans, JPCB_PLAN_MIN=m). */
int jpcb_plan_cache_configure(int max_plans /* 0: disable and drop all plans */, int min_calls);
int jpcb_plan_cache_clear(void);
void jpcb_plan_cache_stats(unsigned long long* replays, unsigned long long* captures,
                           unsigned long long* bypassed);

/* ---- positional form of the compute entry points (csrc/call.cu) -----------------------------------------------------
 * What an XLA custom call has in hand (jax.ffi -- the binding a jpc maintainer would add; the reference's
 * functions are `eqx.filter_jit`ed, jpc/_train.py:35, so under JAX every call arrives this way): ONE flat list of operand
 * buffers, ONE flat list of result buffers, the descriptor as an opaque byte attribute.  jpcb_call splits the lists into
 * the pointer lists of the structured entry point `entry` and calls it; jpcb_call_counts says how many buffers the
 * descriptor implies (host arithmetic only).  The LAST result is always the workspace (`workspace_bytes` long).  With
 *   L = n_layers, [b] present iff has_bias, S = the state list (L entries; L + 1 when free_input), [x] absent iff free_input,
 *   bPC: n = td.n_layers entries per list, [bW] iff td.has_bias, [bV] iff bu_has_bias:
 *   PC_FFWD_INIT      args W [b] x y          rets A(L) loss ws
 *   PC_ENERGY         args W [b] S [x] y      rets E_layers ws
 *   PC_ACTIVITY_GRAD  args W [b] S [x] y      rets F gS ws
 *   PC_INFER          args W [b] S [x] y      rets S' ws          (S' may be the S buffers themselves -- donated operands,
 *                                                                 input_output_aliases -- else S is copied into S' first)
 *   PC_PARAM_GRADS    args W [b] S [x] y      rets gW [gb] E_layers ws
 *   PC_TRAIN_STEP     args W [b] x y          rets A(L) loss E_layers gW [gb] ws
 *   HPC_PARAM_GRADS   args W [b] U(L) T(L)    rets gW [gb] E_layers ws
 *   EPC_GRADS         args W [b] S [x] y      rets F gS gW [gb] ws
 *   BPC_ENERGY        args W [bW] V [bV] A [x] y   rets E_pairs ws
 *   BPC_ACTIVITY_GRAD args (same)                  rets F gA ws
 *   BPC_INFER         args (same)                  rets A' ws
 *   BPC_PARAM_GRADS   args (same)                  rets gW [gbW] gV [gbV] E_pairs ws
 *   BPC_TRAIN_STEP    args (same)                  rets A' gW [gbW] gV [gbV] E_pairs ws
 * `desc` is a jpcb_pc_desc (jpcb_bpc_desc for the BPC entries); desc_bytes must be its sizeof. */
enum {
  JPCB_CALL_PC_FFWD_INIT = 0,
  JPCB_CALL_PC_ENERGY = 1,
  JPCB_CALL_PC_ACTIVITY_GRAD = 2,
  JPCB_CALL_PC_INFER = 3,
  JPCB_CALL_PC_PARAM_GRADS = 4,
  JPCB_CALL_PC_TRAIN_STEP = 5,
  JPCB_CALL_HPC_PARAM_GRADS = 6,
  JPCB_CALL_EPC_GRADS = 7,
  JPCB_CALL_BPC_ENERGY = 8,
  JPCB_CALL_BPC_ACTIVITY_GRAD = 9,
  JPCB_CALL_BPC_INFER = 10,
  JPCB_CALL_BPC_PARAM_GRADS = 11,
  JPCB_CALL_BPC_TRAIN_STEP = 12,
  JPCB_CALL_N_ENTRIES = 13
};
int jpcb_call_counts(int entry, const void* desc, size_t desc_bytes, int* n_args, int* n_rets);
int jpcb_call(int entry, const void* desc, size_t desc_bytes, const void* const* args, int n_args,
              void* const* rets, int n_rets, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JPC_B200_H_ */
