// bpc.cu -- bidirectional predictive coding (reference: jpc/_core/_energies.py:246-357 `bpc_energy_fn`,
// jpc/_core/_grads.py:170-226,544-612, jpc/_core/_updates.py:206-372) on the grouped-GEMM kernels.
//
// Index map (H = n_layers-1 of the top-down model; D[0..H+1] = td.dims; activities A[0..H-1] =
// z_0..z_{H-1}, A[H] = unused dummy slot):
//   top-down  layer j = 0..H : W_j [D[j+1], D[j]],  input u_j = x (j=0) | z_{j-1},  target t_j = z_j | y (j=H)
//       e_j     = t_j - (W_j phi_j(u_j) + b_j) - s_j u_j              (s_j = 1 only for 1 <= j <= H-1)
//   bottom-up layer j = 0..H : V_j [D[j], D[j+1]],  input v_j = z_j | y (j=H),      target tau_j = x (j=0) | z_{j-1}
//       delta_j = tau_j - (V_j psi_j(v_j) + c_j) - s_j v_j            (s_j = 1 only for 1 <= j <= H-1)
//   F = (alpha_f sum_j |e_j|^2/2 + alpha_b sum_j |delta_j|^2/2) / B     (no param_type scalings, _energies.py:314)
//   dF/dz_k = (1/B) [ alpha_f (e_k     - phi_{k+1}'(z_k) (e_{k+1} W_{k+1}) - s_{k+1} e_{k+1})
//                   + alpha_b (delta_{k+1} - psi_k'(z_k)   (delta_k V_k)     - s_k delta_k) ]
//   dF/dW_j = -(alpha_f/B) e_j^T phi_j(u_j)        dF/dV_j = -(alpha_b/B) delta_j^T psi_j(v_j)
// One evaluation = three grouped launches: all 2(H+1) errors; the forward-direction half of every
// activity gradient (into a scratch buffer); the backward-direction half + combination + Euler
// update.  The two constant predictions (W_0 phi_0(x) and V_H psi_H(y)) are hoisted out of the loop.
//
// td.dtype selects the arithmetic of the GEMM operands exactly as on the PC path: JPCB_DTYPE_F32 = CUDA-core
// fp32 FMA (1e-5 parity); JPCB_DTYPE_BF16 = the tcgen05 grouped GEMM on bf16 operand copies with fp32
// accumulation and fp32 master activities (2e-2 parity).  The bf16 path keeps ONE operand copy bf16(act(z_k))
// per activity, so it needs the two directions to apply the same activation to z_k
// (td.act[k+1] == bu_act[k], true for make_mlp models), and every dim / the batch to be multiples of 8
// (the host mirror zero-pads feature dims).
#include <cuda_runtime.h>

#include <cstdlib>
#include <vector>

#include "elementwise.cuh"
#include "host.cuh"

using namespace jpcb;
typedef __nv_bfloat16 bf16;

namespace {

// acc[0..H] = sum |e_j|^2/2, acc[H+1..2H+1] = sum |delta_j|^2/2  ->  reference pair order
// [(e_H, delta_0), (e_0, delta_1), ..., (e_{H-1}, delta_H)] / B  (jpc/_core/_energies.py:322-355)
__global__ void bpc_finalize_kernel(const float* __restrict__ acc, int H, float inv_b, float af, float ab, float* E_pairs, float* F) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float* e = acc;
    const float* d = acc + H + 1;
    float tot = 0.f;
    for (int j = 0; j <= H; ++j) tot += af * e[j] + ab * d[j];
    if (E_pairs) {
      E_pairs[0] = af * e[H] * inv_b;
      E_pairs[1] = ab * d[0] * inv_b;
      for (int l = 0; l < H; ++l) {
        E_pairs[2 + 2 * l] = af * e[l] * inv_b;
        E_pairs[3 + 2 * l] = ab * d[l + 1] * inv_b;
      }
    }
    if (F) *F = tot * inv_b;
  }
}

// x = None (jpc/_core/_energies.py:324: x_eff = activities[0]): z_0 is also the INPUT of top-down layer 0 and the TARGET of
// bottom-up layer 0, which adds  -alpha_f phi_0'(z_0) (e_0 W_0) + alpha_b delta_0  to its gradient.  Gf0 already holds the
// forward-direction half; Gx the first extra term (its own GEMM); delta_0 comes as fp32 or as the bf16 error copy.
__global__ void bpc_free_x_combine_kernel(float* __restrict__ Gf0, const float* __restrict__ Gx, const float* __restrict__ d0,
                                          const __nv_bfloat16* __restrict__ d0b, size_t n, float ab) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float dv = d0 ? d0[i] : __bfloat162float(d0b[i]);
    Gf0[i] += Gx[i] + ab * dv;
  }
}

int validate_bpc(const jpcb_bpc_desc* d) {
  if (!d) return fail(JPCB_EINVAL, "null descriptor");
  const jpcb_pc_desc& t = d->td;
  if (t.n_layers < 2 || t.n_layers > JPCB_MAX_LAYERS) return fail(JPCB_EINVAL, "n_layers=%d outside [2,%d]", t.n_layers, JPCB_MAX_LAYERS);
  if (t.batch_local <= 0 || t.batch_global < t.batch_local) return fail(JPCB_EINVAL, "bad batch sizes local=%d global=%d", t.batch_local, t.batch_global);
  for (int l = 0; l <= t.n_layers; ++l)
    if (t.dims[l] <= 0) return fail(JPCB_EINVAL, "dims[%d]=%d", l, t.dims[l]);
  for (int l = 0; l < t.n_layers; ++l) {
    if (t.act[l] < 0 || t.act[l] > JPCB_ACT_SILU || d->bu_act[l] < 0 || d->bu_act[l] > JPCB_ACT_SILU)
      return fail(JPCB_EINVAL, "Invalid activation function ID at layer %d", l);
    if (t.skip[l] && l >= 1 && l <= t.n_layers - 2 && t.dims[l] != t.dims[l + 1])
      return fail(JPCB_EINVAL, "identity skip at layer %d needs equal dims (%d vs %d)", l, t.dims[l], t.dims[l + 1]);
  }
  if (t.dtype != JPCB_DTYPE_F32 && t.dtype != JPCB_DTYPE_BF16 && t.dtype != JPCB_DTYPE_F32X3) return fail(JPCB_EINVAL, "bad dtype %d", t.dtype);
  if (t.dtype != JPCB_DTYPE_F32) {
    if (t.batch_local % 8) return fail(JPCB_ENOTIMPL, "the tensor-core paths (bf16, f32x3) need batch_local %% 8 == 0 (got %d)", t.batch_local);
    for (int l = 0; l <= t.n_layers; ++l)
      if (t.dims[l] % 8) return fail(JPCB_ENOTIMPL, "the tensor-core paths (bf16, f32x3) need every dim %% 8 == 0 (dims[%d]=%d)", l, t.dims[l]);
    for (int k = 0; k + 1 < t.n_layers; ++k)
      if (t.act[k + 1] != d->bu_act[k])
        return fail(JPCB_ENOTIMPL, "bf16 bPC path needs td.act[%d] == bu_act[%d] (one bf16 operand copy per activity)", k + 1, k);
  }
  for (int l = 0; l <= t.n_layers; ++l)
    if ((long long)t.batch_local * t.dims[l] >= (1LL << 31)) return fail(JPCB_ENOTIMPL, "batch_local * dims[%d] must be < 2^31", l);
  if (t.solver != JPCB_SOLVER_EULER) return fail(JPCB_ENOTIMPL, "bPC inference is Euler (update_bpc_activities with optax.sgd); Heun is not on this path");
  if (t.n_steps < 0) return fail(JPCB_EINVAL, "n_steps=%d", t.n_steps);
  if (t.activity_decay != 0.f) return fail(JPCB_EINVAL, "bpc_energy_fn has no activity_decay");
  if (t.free_input && t.dims[0] != t.dims[1])
    return fail(JPCB_EINVAL, "bPC with x = None uses activities[0] as the input of top_down_model[0] and the target of bottom_up_model[0]: "
                             "needs dims[0] == dims[1] (got %d vs %d)", t.dims[0], t.dims[1]);
  for (int l = 0; l < t.n_layers; ++l)
    if (t.post_act[l] != JPCB_ACT_LINEAR) return fail(JPCB_ENOTIMPL, "bPC: layers of the form act(W u + b) are not on this path");
  return JPCB_OK;
}

struct Bump {
  char* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    off = (off + 1023) & ~(size_t)1023;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct BpcRun {
  const jpcb_bpc_desc* d;
  const float* const* W; const float* const* bW;
  const float* const* V; const float* const* bV;
  const float* x; const float* y;
  cudaStream_t st;
  int H, Bl;
  const int32_t* D;
  float invB, af, ab;
  // workspace
  float* red = nullptr;             // [2(H+1)] energy sums
  std::vector<float*> e, dl, Gf;    // e_j [Bl, D[j+1]], delta_j [Bl, D[j]], forward half of dF/dz_k [Bl, D[k+1]]
  bool fx = false;                  // td.free_input: x = None, x_eff = activities[0] (set by the entry points: x = A[0])
  float* Gx = nullptr;              // fx: -alpha_f phi_0'(z_0) (e_0 W_0)   [Bl, D[0]]
  float* P0 = nullptr;              // W_0 phi_0(x) + b_0        [Bl, D[1]]
  float* QH = nullptr;              // V_H psi_H(y) + c_H        [Bl, D[H]]
  // bf16 operand copies (tensor-core paths).  x3: the fp32-parity path -- every copy is a 3-way split [rows, 3 P(d)]
  // (hi | mid | lo, P(d) = d rounded up to the 64-wide k-block, padding zero) and every GEMM runs six passes.
  bool tc = false, x3 = false;
  int P(int dd) const { return x3 ? (dd + 63) / 64 * 64 : dd; }
  size_t cw(int dd) const { return x3 ? 3 * (size_t)P(dd) : (size_t)dd; }
  void set_dtype(int dtype) { tc = dtype != JPCB_DTYPE_F32; x3 = dtype == JPCB_DTYPE_F32X3; }
  // operand geometry of a tensor-core task whose concatenated dimension is `ka` (A) / `kb` (B) wide
  void split_geom(GTask& t, int ka, int kb) const {
    if (!x3) return;
    static const int passes = [] { const char* v = getenv("JPCB_X3_PASSES"); const int n = v ? atoi(v) : 6; return n >= 3 && n <= 6 ? n : 6; }();
    t.split = passes; t.a_ld = (int)cw(ka); t.b_ld = (int)cw(kb); t.a_part = P(ka); t.b_part = P(kb);
  }
  void set_ob(GTask& t, bf16* p) const { t.e.Ob = p; if (x3 && p) { t.e.ob_ld = (int)cw(t.e.N); t.e.ob_part = P(t.e.N); } }
  std::vector<bf16*> Wb, WTb, Vb, VTb;  // W_j, W_j^T (j >= 1), V_j, V_j^T (j <= H-1)
  std::vector<bf16*> Hz;                // act(z_k)      [Bl, D[k+1]], k = 0..H-1
  bf16* Hx = nullptr;                   // phi_0(x)      [Bl, D[0]]
  bf16* Hy = nullptr;                   // psi_H(y)      [Bl, D[H+1]]
  std::vector<bf16*> Eb, Db;            // bf16 copies of e_j / delta_j
  size_t bytes = 0;

  void carve(void* base) {
    Bump b{reinterpret_cast<char*>(base)};
    red = b.take<float>(2 * (H + 1) + 8);
    e.assign(H + 1, nullptr); dl.assign(H + 1, nullptr); Gf.resize(H);
    Wb.assign(H + 1, nullptr); WTb.assign(H + 1, nullptr); Vb.assign(H + 1, nullptr); VTb.assign(H + 1, nullptr);
    Hz.assign(H, nullptr); Eb.assign(H + 1, nullptr); Db.assign(H + 1, nullptr);
    if (!tc || x3) {  // fp32 errors (the split path keeps them too: e_l / skip sources of the gradient epilogue, bias gradients)
      for (int j = 0; j <= H; ++j) e[j] = b.take<float>((size_t)Bl * D[j + 1]);
      for (int j = 0; j <= H; ++j) dl[j] = b.take<float>((size_t)Bl * D[j]);
    }
    if (tc) {
      for (int j = 0; j <= H; ++j) {
        Wb[j] = b.take<bf16>((size_t)D[j + 1] * cw(D[j])); Vb[j] = b.take<bf16>((size_t)D[j] * cw(D[j + 1]));
        if (j >= 1 || fx) WTb[j] = b.take<bf16>((size_t)D[j] * cw(D[j + 1]));
        if (j <= H - 1) VTb[j] = b.take<bf16>((size_t)D[j + 1] * cw(D[j]));
        Eb[j] = b.take<bf16>((size_t)Bl * cw(D[j + 1]));
        Db[j] = b.take<bf16>((size_t)Bl * cw(D[j]));
      }
      for (int k = 0; k < H; ++k) Hz[k] = b.take<bf16>((size_t)Bl * cw(D[k + 1]));
      Hx = b.take<bf16>((size_t)Bl * cw(D[0]));
      Hy = b.take<bf16>((size_t)Bl * cw(D[H + 1]));
    }
    for (int k = 0; k < H; ++k) Gf[k] = b.take<float>((size_t)Bl * D[k + 1]);
    if (fx) Gx = b.take<float>((size_t)Bl * D[0]);
    P0 = b.take<float>((size_t)Bl * D[1]);
    QH = b.take<float>((size_t)Bl * D[H]);
    bytes = ((b.off + 1023) & ~(size_t)1023) + 1024;
  }

  int init(const jpcb_bpc_desc* desc, const float* const* W_, const float* const* bW_, const float* const* V_, const float* const* bV_,
           const float* x_, const float* y_, void* ws, size_t ws_bytes, void* stream) {
    RET(validate_bpc(desc));
    d = desc; W = W_; bW = bW_; V = V_; bV = bV_; x = x_; y = y_;
    st = reinterpret_cast<cudaStream_t>(stream);
    H = d->td.n_layers - 1; Bl = d->td.batch_local; D = d->td.dims;
    invB = 1.f / (float)d->td.batch_global; af = d->forward_energy_weight; ab = d->backward_energy_weight;
    set_dtype(d->td.dtype);
    fx = d->td.free_input != 0;
    if (fx && x) return fail(JPCB_EINVAL, "free_input (x = None): `x` must be NULL, activities[0] takes its place");
    if (!W || !V || (!x && !fx) || !y) return fail(JPCB_EINVAL, "null weight lists / input / target");
    if (d->td.has_bias && !bW) return fail(JPCB_EINVAL, "top-down has_bias set but no bias list");
    if (d->bu_has_bias && !bV) return fail(JPCB_EINVAL, "bottom-up has_bias set but no bias list");
    carve(ws);
    if (!ws || ws_bytes < bytes) return fail(JPCB_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", bytes, ws_bytes);
    return JPCB_OK;
  }
  float skip(int j) const { return (d->td.skip[j] && j >= 1 && j <= H - 1) ? 1.f : 0.f; }
  const float* biasW(int j) const { return d->td.has_bias ? bW[j] : nullptr; }
  const float* biasV(int j) const { return d->bu_has_bias ? bV[j] : nullptr; }

  int launch(const std::vector<GTask>& t) { return tc ? launch_tc(t, st) : launch_simt(t, st); }
  static int ew_grid(size_t items, int per_block) {
    size_t blocks = (items + per_block - 1) / per_block;
    const size_t cap = (size_t)num_sms() * 8;
    return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
  }
  // bf16 operand copies of the weights (both directions, plus the transposes the gradient GEMMs read) and of
  // the clamped inputs; with `S` also of every activity
  int split3(const float* src, bf16* dst, int R, int C, int act) {
    split3_kernel<<<ew_grid((size_t)R * P(C), 256), 256, 0, st>>>(src, dst, R, C, P(C), act); count_launch();
    return JPCB_OK;
  }
  int tsplit3(const float* src, bf16* dst, int R, int C) {  // src [R, C] -> split3(src^T) [C, 3 P(R)]
    transpose_split3_kernel<<<ew_grid(((P(R) + 31) / 32) * ((C + 31) / 32), 1), dim3(32, 8), 0, st>>>(src, dst, R, C, P(R)); count_launch();
    return JPCB_OK;
  }
  int prep_x3(const float* const* S) {
    for (int j = 0; j <= H; ++j) {
      split3(W[j], Wb[j], D[j + 1], D[j], ACT_LINEAR);
      split3(V[j], Vb[j], D[j], D[j + 1], ACT_LINEAR);
      if (j >= 1 || fx) tsplit3(W[j], WTb[j], D[j + 1], D[j]);
      if (j <= H - 1) tsplit3(V[j], VTb[j], D[j], D[j + 1]);
      // error copies are written by the epilogues (valid columns only): their padding columns must read as zero
      CK(cudaMemsetAsync(Eb[j], 0, (size_t)Bl * cw(D[j + 1]) * sizeof(bf16), st));
      CK(cudaMemsetAsync(Db[j], 0, (size_t)Bl * cw(D[j]) * sizeof(bf16), st));
    }
    split3(x, Hx, Bl, D[0], d->td.act[0]);
    split3(y, Hy, Bl, D[H + 1], d->bu_act[H]);
    if (S)
      for (int k = 0; k < H; ++k) split3(S[k], Hz[k], Bl, D[k + 1], d->td.act[k + 1]);
    CK(cudaGetLastError());
    return JPCB_OK;
  }
  int prep(const float* const* S) {
    if (!tc) return JPCB_OK;
    if (x3) return prep_x3(S);
    {  // every weight of both models read once: bf16 copy + transposed bf16 copy where a transposed GEMM needs it
      PrepJobs jobs;
      jobs.n = 0; jobs.total = 0;
      auto flush = [&]() {
        if (jobs.n == 0) return;
        prep_weights_kernel<<<ew_grid(jobs.total, 1), 256, 0, st>>>(jobs); count_launch();
        jobs.n = 0; jobs.total = 0;
      };
      auto add = [&](const float* Wp, bf16* Wbp, bf16* WTp, int R, int C) {
        if (jobs.n == PREP_MAX_JOBS) flush();
        PrepJob& jb = jobs.j[jobs.n++];
        jb.W = Wp; jb.Wb = Wbp; jb.WTb = WTp; jb.R = R; jb.C = C;
        jb.tile_begin = jobs.total; jb.tiles_c = (C + 63) / 64;
        jobs.total += jb.tiles_c * ((R + 63) / 64);
      };
      for (int j = 0; j <= H; ++j) {
        add(W[j], Wb[j], (j >= 1 || fx) ? WTb[j] : nullptr, D[j + 1], D[j]);      // W_j [D[j+1], D[j]]
        add(V[j], Vb[j], j <= H - 1 ? VTb[j] : nullptr, D[j], D[j + 1]);  // V_j [D[j], D[j+1]]
      }
      flush();
    }
    const size_t nx = (size_t)Bl * D[0], ny = (size_t)Bl * D[H + 1];
    cast_act_bf16_kernel<<<ew_grid(nx / 4 + 1, 256), 256, 0, st>>>(x, Hx, nx, d->td.act[0]); count_launch();
    cast_act_bf16_kernel<<<ew_grid(ny / 4 + 1, 256), 256, 0, st>>>(y, Hy, ny, d->bu_act[H]); count_launch();
    if (S)
      for (int k = 0; k < H; ++k) {
        const size_t n = (size_t)Bl * D[k + 1];
        cast_act_bf16_kernel<<<ew_grid(n / 4 + 1, 256), 256, 0, st>>>(S[k], Hz[k], n, d->td.act[k + 1]); count_launch();
      }
    CK(cudaGetLastError());
    return JPCB_OK;
  }

  // top-down prediction GEMM of layer j on state S: phi_j(u_j) W_j^T
  void td_operands(GTask& t, int j, const float* const* S) const {
    t.e.M = Bl; t.e.N = D[j + 1]; t.K = D[j];
    t.a_p = j == 0 ? x : S[j - 1]; t.a_smn = D[j]; t.a_sk = 1; t.a_act = d->td.act[j];
    t.b_p = W[j]; t.b_smn = D[j]; t.b_sk = 1;
    t.a_b = j == 0 ? Hx : Hz[j - 1]; t.b_b = Wb[j];
    split_geom(t, D[j], D[j]);
    t.e.alpha = 1.f; t.e.bias = biasW(j);
  }
  // bottom-up prediction GEMM of layer j on state S: psi_j(v_j) V_j^T
  void bu_operands(GTask& t, int j, const float* const* S) const {
    t.e.M = Bl; t.e.N = D[j]; t.K = D[j + 1];
    t.a_p = j == H ? y : S[j]; t.a_smn = D[j + 1]; t.a_sk = 1; t.a_act = d->bu_act[j];
    t.b_p = V[j]; t.b_smn = D[j + 1]; t.b_sk = 1;
    t.a_b = j == H ? Hy : Hz[j]; t.b_b = Vb[j];
    split_geom(t, D[j + 1], D[j + 1]);
    t.e.alpha = 1.f; t.e.bias = biasV(j);
  }

  // the two predictions whose inputs are clamped: P0 = f_0(x), QH = g_H(y)
  int hoist() {
    GTask a, b;
    td_operands(a, 0, nullptr); a.e.mode = EPI_FFWD; a.e.O = P0;
    bu_operands(b, H, nullptr); b.e.mode = EPI_FFWD; b.e.O = QH;
    return launch({a, b});
  }

  // all 2(H+1) prediction errors on state S (+ energy sums when `energy`)
  int errors(const float* const* S, bool energy, bool hoisted) {
    std::vector<GTask> ts, k0;  // k0: the two GEMM-free tasks (always on the CUDA-core kernel)
    for (int j = 0; j <= H; ++j) {
      GTask t;
      t.e.mode = EPI_ERR; t.e.escale = 1.f;
      if (j == 0 && hoisted && !fx) {  // e_0 = z_0 - P0 without a GEMM
        t.e.M = Bl; t.e.N = D[1]; t.K = 0; t.e.alpha = 0.f; t.e.skip = 1.f; t.e.U = P0;
        t.a_p = S[0]; t.b_p = S[0];
      } else {
        td_operands(t, j, S);
        t.e.skip = skip(j); t.e.U = j == 0 ? x : S[j - 1];
      }
      t.e.T = j == H ? y : S[j];
      t.e.O = e[j]; set_ob(t, Eb[j]);
      if (energy) t.e.red = red + j;
      (t.K == 0 ? k0 : ts).push_back(t);
    }
    for (int j = 0; j <= H; ++j) {
      GTask t;
      t.e.mode = EPI_ERR; t.e.escale = 1.f;
      if (j == H && hoisted) {  // delta_H = z_{H-1} - QH
        t.e.M = Bl; t.e.N = D[H]; t.K = 0; t.e.alpha = 0.f; t.e.skip = 1.f; t.e.U = QH;
        t.a_p = S[0]; t.b_p = S[0];
      } else {
        bu_operands(t, j, S);
        t.e.skip = skip(j); t.e.U = j == H ? y : S[j];
      }
      t.e.T = j == 0 ? x : S[j - 1];
      t.e.O = dl[j]; set_ob(t, Db[j]);
      if (energy) t.e.red = red + H + 1 + j;
      (t.K == 0 ? k0 : ts).push_back(t);
    }
    // the two GEMM-free errors: streaming kernel when the layout allows 16-byte vectors, else the CUDA-core kernel's K = 0 path
    if (hoisted && !fx && D[1] % 4 == 0 && D[H] % 4 == 0 && (reinterpret_cast<uintptr_t>(S[0]) & 15) == 0 && (reinterpret_cast<uintptr_t>(S[H - 1]) & 15) == 0) {
      RET(launch(ts));
      const size_t n0 = (size_t)Bl * D[1] / 4, nH = (size_t)Bl * D[H] / 4;
      const Err0Job j0{S[0], P0, e[0], Eb[0], n0, energy ? red + 0 : nullptr, D[1], x3 ? (int)cw(D[1]) : 0, x3 ? P(D[1]) : 0};
      const Err0Job jH{S[H - 1], QH, dl[H], Db[H], nH, energy ? red + H + 1 + H : nullptr, D[H], x3 ? (int)cw(D[H]) : 0, x3 ? P(D[H]) : 0};
      err0_pair_kernel<<<dim3(ew_grid((n0 > nH ? n0 : nH), 256), 2), 256, 0, st>>>(j0, jH);
      count_launch();
      CK(cudaGetLastError());
      return JPCB_OK;
    }
    if (!tc) { ts.insert(ts.end(), k0.begin(), k0.end()); return launch_simt(ts, st); }
    RET(launch_tc(ts, st));
    return launch_simt(k0, st);
  }

  // dF/dz_k (sub = GRAD_OUT, c = 1/B, into Out[k]) or the Euler update z_k -= c * B dF/dz_k (in place)
  int grads(int sub, float c, const float* const* Z, float* const* Out) {
    std::vector<GTask> f, b;
    for (int k = 0; k < H; ++k) {
      GTask t;  // forward-direction half: alpha_f (e_k - phi_{k+1}'(z_k) (e_{k+1} W_{k+1}) - s_{k+1} e_{k+1})
      t.e.mode = EPI_GRAD; t.e.sub = GRAD_OUT; t.e.act = d->td.act[k + 1];
      t.e.M = Bl; t.e.N = D[k + 1]; t.K = D[k + 2];
      t.a_p = e[k + 1]; t.a_smn = D[k + 2]; t.a_sk = 1;
      t.b_p = W[k + 1]; t.b_smn = 1; t.b_sk = D[k + 1];  // B(n,k') = W_{k+1}[k'][n]
      t.a_b = Eb[k + 1]; t.b_b = WTb[k + 1];
      split_geom(t, D[k + 2], D[k + 2]);
      // (split path: the bf16 error copies are 3-part operand copies, not readable as plain bf16 -- fp32 errors instead)
      t.e.alpha = 1.f; t.e.skip = skip(k + 1); t.e.En = e[k + 1]; t.e.Enb = x3 ? nullptr : Eb[k + 1];
      t.e.Z = Z[k]; t.e.El = e[k]; t.e.Elb = x3 ? nullptr : Eb[k];
      t.e.gscale = af; t.e.c = 1.f; t.e.O = Gf[k];
      f.push_back(t);
      GTask u;  // backward-direction half + combination (+ update)
      u.e.mode = EPI_GRAD; u.e.sub = sub; u.e.act = d->bu_act[k];
      u.e.M = Bl; u.e.N = D[k + 1]; u.K = D[k];
      u.a_p = dl[k]; u.a_smn = D[k]; u.a_sk = 1;
      u.b_p = V[k]; u.b_smn = 1; u.b_sk = D[k + 1];       // B(n,k') = V_k[k'][n]
      u.a_b = Db[k]; u.b_b = VTb[k];
      split_geom(u, D[k], D[k]);
      u.e.alpha = 1.f; u.e.skip = skip(k); u.e.En = dl[k]; u.e.Enb = x3 ? nullptr : Db[k];
      u.e.Z = Z[k]; u.e.El = dl[k + 1]; u.e.Elb = x3 ? nullptr : Db[k + 1];
      u.e.gscale = ab; u.e.Gin = Gf[k]; u.e.c = c; u.e.O = Out[k];
      if (tc && sub != GRAD_OUT) set_ob(u, Hz[k]);  // act(z_new): the next evaluation's operand (both directions)
      b.push_back(u);
    }
    if (fx) {  // x = None: the term z_0 receives as the INPUT of top-down layer 0 (own GEMM), see bpc_free_x_combine_kernel
      GTask t;
      t.e.mode = EPI_GRAD; t.e.sub = GRAD_OUT; t.e.act = d->td.act[0];
      t.e.M = Bl; t.e.N = D[0]; t.K = D[1];
      t.a_p = e[0]; t.a_smn = D[1]; t.a_sk = 1;
      t.b_p = W[0]; t.b_smn = 1; t.b_sk = D[0];           // B(n,k') = W_0[k'][n]
      t.a_b = Eb[0]; t.b_b = WTb[0];
      split_geom(t, D[1], D[1]);
      t.e.alpha = 1.f; t.e.Z = Z[0]; t.e.P = Z[0];         // "own error" zero: e = Z - P
      t.e.gscale = af; t.e.c = 1.f; t.e.O = Gx;
      f.push_back(t);
    }
    RET(launch(f));
    if (fx) {
      const size_t n0 = (size_t)Bl * D[0];
      bpc_free_x_combine_kernel<<<ew_grid(n0, 256), 256, 0, st>>>(Gf[0], Gx, dl[0], dl[0] ? nullptr : Db[0], n0, ab);
      count_launch();
      CK(cudaGetLastError());
    }
    RET(launch(b));
    if (fx && tc && sub != GRAD_OUT) {  // z_0 moved: refresh the operand copy phi_0(z_0) that top-down layer 0 reads
      if (x3) split3(Z[0], Hx, Bl, D[0], d->td.act[0]);
      else {
        const size_t nx = (size_t)Bl * D[0];
        cast_act_bf16_kernel<<<ew_grid(nx / 4 + 1, 256), 256, 0, st>>>(Z[0], Hx, nx, d->td.act[0]); count_launch();
      }
      CK(cudaGetLastError());
    }
    return JPCB_OK;
  }

  // both models' parameter gradients from the errors of the last errors() call on state A (jpc/_core/_grads.py:544-612)
  int wgrads(const float* const* A, float* const* gW, float* const* gbW, float* const* gV, float* const* gbV) {
    std::vector<GTask> ts;
    for (int j = 0; j <= H; ++j) {
      GTask t;  // gW_j = -(alpha_f/B) e_j^T phi_j(u_j)
      t.e.mode = EPI_SCALE; t.e.M = D[j + 1]; t.e.N = D[j]; t.K = Bl;
      t.a_p = e[j]; t.a_smn = 1; t.a_sk = D[j + 1];
      t.b_p = j == 0 ? x : A[j - 1]; t.b_smn = 1; t.b_sk = D[j]; t.b_act = d->td.act[j];
      t.a_b = Eb[j]; t.b_b = j == 0 ? Hx : Hz[j - 1]; t.mn_major = tc;  // both read MN-major: no transposes
      split_geom(t, D[j + 1], D[j]);
      t.e.alpha = -af * invB; t.e.O = gW[j];
      ts.push_back(t);
      GTask u;  // gV_j = -(alpha_b/B) delta_j^T psi_j(v_j)
      u.e.mode = EPI_SCALE; u.e.M = D[j]; u.e.N = D[j + 1]; u.K = Bl;
      u.a_p = dl[j]; u.a_smn = 1; u.a_sk = D[j];
      u.b_p = j == H ? y : A[j]; u.b_smn = 1; u.b_sk = D[j + 1]; u.b_act = d->bu_act[j];
      u.a_b = Db[j]; u.b_b = j == H ? Hy : Hz[j]; u.mn_major = tc;
      split_geom(u, D[j], D[j + 1]);
      u.e.alpha = -ab * invB; u.e.O = gV[j];
      ts.push_back(u);
    }
    RET(launch(ts));
    if (d->td.has_bias) {
      if (!gbW) return fail(JPCB_EINVAL, "top-down has_bias set but no bias-gradient list");
      for (int j = 0; j <= H; ++j) {
        if (tc && !x3) colsum_kernel<bf16><<<(D[j + 1] + 31) / 32, dim3(32, 8), 0, st>>>(Eb[j], gbW[j], Bl, D[j + 1], -af * invB);
        else colsum_kernel<float><<<(D[j + 1] + 31) / 32, dim3(32, 8), 0, st>>>(e[j], gbW[j], Bl, D[j + 1], -af * invB);
        count_launch();
      }
    }
    if (d->bu_has_bias) {
      if (!gbV) return fail(JPCB_EINVAL, "bottom-up has_bias set but no bias-gradient list");
      for (int j = 0; j <= H; ++j) {
        if (tc && !x3) colsum_kernel<bf16><<<(D[j] + 31) / 32, dim3(32, 8), 0, st>>>(Db[j], gbV[j], Bl, D[j], -ab * invB);
        else colsum_kernel<float><<<(D[j] + 31) / 32, dim3(32, 8), 0, st>>>(dl[j], gbV[j], Bl, D[j], -ab * invB);
        count_launch();
      }
    }
    CK(cudaGetLastError());
    return JPCB_OK;
  }

  int finalize(float* E_pairs, float* F) {
    bpc_finalize_kernel<<<1, 32, 0, st>>>(red, H, invB, af, ab, E_pairs, F);
    count_launch();
    CK(cudaGetLastError());
    return JPCB_OK;
  }
  int zero_red() { CK(cudaMemsetAsync(red, 0, (2 * (H + 1) + 8) * sizeof(float), st)); return JPCB_OK; }
};

}  // namespace

extern "C" {

size_t jpcb_bpc_workspace_bytes(const jpcb_bpc_desc* desc) {
  if (validate_bpc(desc) != JPCB_OK) return 0;
  BpcRun r;
  r.d = desc; r.H = desc->td.n_layers - 1; r.Bl = desc->td.batch_local; r.D = desc->td.dims;
  r.set_dtype(desc->td.dtype);
  r.fx = desc->td.free_input != 0;
  r.carve(nullptr);
  return r.bytes;
}

static int bpc_energy_impl(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW, const float* const* V,
                    const float* const* bV, const float* const* A, const float* x, const float* y, float* E_pairs, void* workspace,
                    size_t workspace_bytes, void* stream) {
  BpcRun r;
  RET(r.init(desc, W, bW, V, bV, x, y, workspace, workspace_bytes, stream));
  if (!A || !E_pairs) return fail(JPCB_EINVAL, "null activities / output");
  if (r.fx) r.x = A[0];
  RET(r.zero_red());
  RET(r.prep(A));
  RET(r.errors(A, true, false));
  return r.finalize(E_pairs, nullptr);
}

static int bpc_activity_grad_impl(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW, const float* const* V,
                           const float* const* bV, const float* const* A, const float* x, const float* y, float* F, float* const* gA,
                           void* workspace, size_t workspace_bytes, void* stream) {
  BpcRun r;
  RET(r.init(desc, W, bW, V, bV, x, y, workspace, workspace_bytes, stream));
  if (!A || !gA) return fail(JPCB_EINVAL, "null activities / output");
  if (r.fx) r.x = A[0];
  RET(r.zero_red());
  RET(r.prep(A));
  RET(r.errors(A, true, false));
  RET(r.grads(GRAD_OUT, r.invB, A, gA));
  CK(cudaMemsetAsync(gA[r.H], 0, (size_t)r.Bl * r.D[r.H + 1] * sizeof(float), r.st));  // dummy slot: dF/dA[-1] = 0
  if (F) RET(r.finalize(nullptr, F));
  return JPCB_OK;
}

static int bpc_infer_impl(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW, const float* const* V,
                   const float* const* bV, float* const* A, const float* x, const float* y, void* workspace, size_t workspace_bytes,
                   void* stream) {
  BpcRun r;
  RET(r.init(desc, W, bW, V, bV, x, y, workspace, workspace_bytes, stream));
  if (!A) return fail(JPCB_EINVAL, "null activities");
  if (r.fx) r.x = A[0];
  RET(r.prep(A));
  if (!r.fx) RET(r.hoist());   // (x = None: neither clamped-side prediction is hoisted -- P0 is not constant)
  const int T = desc->td.n_steps;
  for (int s = 0; s < T; ++s) {
    const float h = s == T - 1 ? desc->td.dt_last : desc->td.dt;
    RET(r.errors(A, false, !r.fx));
    RET(r.grads(GRAD_EULER, h * r.invB, A, A));
  }
  return JPCB_OK;
}

static int bpc_param_grads_impl(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW, const float* const* V,
                         const float* const* bV, const float* const* A, const float* x, const float* y, float* const* gW,
                         float* const* gbW, float* const* gV, float* const* gbV, float* E_pairs, void* workspace, size_t workspace_bytes,
                         void* stream) {
  BpcRun r;
  RET(r.init(desc, W, bW, V, bV, x, y, workspace, workspace_bytes, stream));
  if (!A || !gW || !gV) return fail(JPCB_EINVAL, "null activities / outputs");
  if (r.fx) r.x = A[0];
  RET(r.zero_red());
  RET(r.prep(A));
  RET(r.errors(A, E_pairs != nullptr, false));
  RET(r.wgrads(A, gW, gbW, gV, gbV));
  CK(cudaGetLastError());
  if (E_pairs) RET(r.finalize(E_pairs, nullptr));
  return JPCB_OK;
}


// solve + parameter gradients in one call: operand preparation once, the activities' operand copies written by the last
// Euler step's epilogues are the weight-gradient GEMMs' operands
static int bpc_train_step_impl(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW, const float* const* V,
                               const float* const* bV, float* const* A, const float* x, const float* y, float* const* gW,
                               float* const* gbW, float* const* gV, float* const* gbV, float* E_pairs, void* workspace,
                               size_t workspace_bytes, void* stream) {
  BpcRun r;
  RET(r.init(desc, W, bW, V, bV, x, y, workspace, workspace_bytes, stream));
  if (!A || !gW || !gV) return fail(JPCB_EINVAL, "null activities / outputs");
  if (r.fx) r.x = A[0];
  RET(r.zero_red());
  RET(r.prep(A));
  if (!r.fx) RET(r.hoist());
  const int T = desc->td.n_steps;
  for (int s = 0; s < T; ++s) {
    const float h = s == T - 1 ? desc->td.dt_last : desc->td.dt;
    RET(r.errors(A, false, !r.fx));
    RET(r.grads(GRAD_EULER, h * r.invB, A, A));
  }
  RET(r.errors(A, E_pairs != nullptr, !r.fx));   // errors (+ energies) at the solution; the two clamped-side ones without a GEMM
  RET(r.wgrads(A, gW, gbW, gV, gbV));
  if (E_pairs) RET(r.finalize(E_pairs, nullptr));
  return JPCB_OK;
}

// ---- planned entry points: identical calls are captured into a CUDA graph once and replayed (plan.cu) ----
namespace {
inline bool key_bpc(PlanKey& k, const jpcb_bpc_desc* d, const float* const* W, const float* const* bW, const float* const* V,
                    const float* const* bV, const float* const* A, const float* x, const float* y, const void* ws, size_t ws_bytes) {
  if (!d || d->td.n_layers < 2 || d->td.n_layers > JPCB_MAX_LAYERS || !W || !V || !A) return false;
  const int n = d->td.n_layers;
  k.bytes(d, sizeof(*d));
  k.list(W, n); k.list(d->td.has_bias ? bW : nullptr, n); k.list(V, n); k.list(d->bu_has_bias ? bV : nullptr, n);
  k.list(A, n); k.ptr(x); k.ptr(y); k.ptr(ws); k.word(ws_bytes);
  return true;
}
}  // namespace

int jpcb_bpc_energy(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW, const float* const* V,
                    const float* const* bV, const float* const* A, const float* x, const float* y, float* E_pairs, void* workspace,
                    size_t workspace_bytes, void* stream) {
  auto body = [&](void* s) { return bpc_energy_impl(desc, W, bW, V, bV, A, x, y, E_pairs, workspace, workspace_bytes, s); };
  PlanKey k(0x7101);
  if (!key_bpc(k, desc, W, bW, V, bV, A, x, y, workspace, workspace_bytes)) return body(stream);
  k.ptr(E_pairs);
  return plan_run(k, stream, body);
}

int jpcb_bpc_activity_grad(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW, const float* const* V,
                           const float* const* bV, const float* const* A, const float* x, const float* y, float* F, float* const* gA,
                           void* workspace, size_t workspace_bytes, void* stream) {
  auto body = [&](void* s) { return bpc_activity_grad_impl(desc, W, bW, V, bV, A, x, y, F, gA, workspace, workspace_bytes, s); };
  PlanKey k(0x7102);
  if (!gA || !key_bpc(k, desc, W, bW, V, bV, A, x, y, workspace, workspace_bytes)) return body(stream);
  k.ptr(F); k.list(gA, desc->td.n_layers);
  return plan_run(k, stream, body);
}

int jpcb_bpc_infer(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW, const float* const* V,
                   const float* const* bV, float* const* A, const float* x, const float* y, void* workspace, size_t workspace_bytes,
                   void* stream) {
  auto body = [&](void* s) { return bpc_infer_impl(desc, W, bW, V, bV, A, x, y, workspace, workspace_bytes, s); };
  PlanKey k(0x7103);
  if (!key_bpc(k, desc, W, bW, V, bV, A, x, y, workspace, workspace_bytes)) return body(stream);
  return plan_run(k, stream, body);
}

int jpcb_bpc_param_grads(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW, const float* const* V,
                         const float* const* bV, const float* const* A, const float* x, const float* y, float* const* gW,
                         float* const* gbW, float* const* gV, float* const* gbV, float* E_pairs, void* workspace, size_t workspace_bytes,
                         void* stream) {
  auto body = [&](void* s) {
    return bpc_param_grads_impl(desc, W, bW, V, bV, A, x, y, gW, gbW, gV, gbV, E_pairs, workspace, workspace_bytes, s);
  };
  PlanKey k(0x7104);
  if (!gW || !gV || !key_bpc(k, desc, W, bW, V, bV, A, x, y, workspace, workspace_bytes)) return body(stream);
  const int n = desc->td.n_layers;
  k.list(gW, n); k.list(desc->td.has_bias ? gbW : nullptr, n); k.list(gV, n); k.list(desc->bu_has_bias ? gbV : nullptr, n); k.ptr(E_pairs);
  return plan_run(k, stream, body);
}

int jpcb_bpc_train_step(const jpcb_bpc_desc* desc, const float* const* W, const float* const* bW, const float* const* V,
                        const float* const* bV, float* const* A, const float* x, const float* y, float* const* gW,
                        float* const* gbW, float* const* gV, float* const* gbV, float* E_pairs, void* workspace, size_t workspace_bytes,
                        void* stream) {
  auto body = [&](void* s) {
    return bpc_train_step_impl(desc, W, bW, V, bV, A, x, y, gW, gbW, gV, gbV, E_pairs, workspace, workspace_bytes, s);
  };
  PlanKey k(0x7105);
  if (!gW || !gV || !key_bpc(k, desc, W, bW, V, bV, A, x, y, workspace, workspace_bytes)) return body(stream);
  const int n = desc->td.n_layers;
  k.list(gW, n); k.list(desc->td.has_bias ? gbW : nullptr, n); k.list(gV, n); k.list(desc->bu_has_bias ? gbV : nullptr, n); k.ptr(E_pairs);
  return plan_run(k, stream, body);
}

}  // extern "C"
