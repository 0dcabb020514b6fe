// api.cu -- C-ABI (include/jpc_b200.h) of the PC train-step path: validates the descriptor,
// carves the caller's workspace, builds the per-phase grouped GEMM task lists and enqueues the
// kernels on the caller's stream.  No allocation, no synchronisation, no CPU arithmetic.
//
// Index conventions used throughout (reference: jpc/_core/_energies.py:86-126):
//   layer l = 0..L-1 maps its input u_l (u_0 = x, u_l = A[l-1]) to a prediction of its target
//   t_l (t_l = A[l] for l < L-1, t_{L-1} = y);  err[l] = t_l - a_l (W_l act_l(u_l) + b_l) - s_l u_l.
//   A[l-1] (l = 1..L-1) is updated with  g = err[l-1] - act_l'(A[l-1]) (a_l err[l] W_l) - s_l err[l].
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "chain_ffwd.cuh"
#include "elementwise.cuh"
#include "host.cuh"
#include "simt_gemm.cuh"
#include "warp_gemm.cuh"

using namespace jpcb;
typedef __nv_bfloat16 bf16;

namespace {
thread_local std::string g_err;
std::atomic<unsigned long long> g_launches{0};  // kernels enqueued by this library (jpcb_launch_count)
std::atomic<int> g_reproducible{0};             // jpcb_set_reproducible
}  // namespace

namespace jpcb {

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

void count_launch(int n) { g_launches += (unsigned long long)n; }

int num_sms() {
  // per DEVICE (a process may drive several GPUs), cached in a small table; a relaxed race only repeats the query
  static std::atomic<int> cache[64];
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64) {
    v = cache[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
  }
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
  if (dev >= 0 && dev < 64) cache[dev].store(v, std::memory_order_relaxed);
  return v;
}

bool reproducible() { return g_reproducible.load(std::memory_order_relaxed) != 0; }

int tc_flags() {
  static const int f = [] { const char* v = getenv("JPCB_TC_FLAGS"); return v ? atoi(v) : 0; }();
  return f;
}

}  // namespace jpcb

namespace {

// ---------------------------------------------------------------------------------------------
// descriptor validation
// ---------------------------------------------------------------------------------------------
int validate(const jpcb_pc_desc* d, int min_layers = 2) {
  if (!d) return fail(JPCB_EINVAL, "null descriptor");
  if (d->n_layers < min_layers || d->n_layers > JPCB_MAX_LAYERS) return fail(JPCB_EINVAL, "n_layers=%d outside [%d,%d]", d->n_layers, min_layers, JPCB_MAX_LAYERS);
  if (d->batch_local <= 0 || d->batch_global < d->batch_local) return fail(JPCB_EINVAL, "bad batch sizes local=%d global=%d", d->batch_local, d->batch_global);
  for (int l = 0; l <= d->n_layers; ++l)
    if (d->dims[l] <= 0) return fail(JPCB_EINVAL, "dims[%d]=%d", l, d->dims[l]);
  for (int l = 0; l < d->n_layers; ++l) {
    if (d->act[l] < 0 || d->act[l] > JPCB_ACT_SILU) return fail(JPCB_EINVAL, "Invalid activation function ID %d at layer %d", d->act[l], l);
    if (d->skip[l] && d->dims[l] != d->dims[l + 1]) return fail(JPCB_EINVAL, "identity skip at layer %d needs equal dims (%d vs %d)", l, d->dims[l], d->dims[l + 1]);
    if (d->post_act[l] < 0 || d->post_act[l] > JPCB_ACT_SILU) return fail(JPCB_EINVAL, "Invalid activation function ID %d after layer %d", d->post_act[l], l);
    if (d->post_act[l] != JPCB_ACT_LINEAR && d->free_input) return fail(JPCB_ENOTIMPL, "layers of the form act(W u + b) are not on the unsupervised (free_input) path");
  }
  if (d->loss == JPCB_LOSS_CE && d->post_act[d->n_layers - 1] != JPCB_ACT_LINEAR)
    return fail(JPCB_ENOTIMPL, "a cross-entropy output layer of the form act(W u + b) is not on this path (the logits are the layer's output)");
  for (int l = 0; l <= d->n_layers; ++l)
    if ((long long)d->batch_local * d->dims[l] >= (1LL << 31)) return fail(JPCB_ENOTIMPL, "batch_local * dims[%d] must be < 2^31 (32-bit element indices)", l);
  if (d->loss != JPCB_LOSS_MSE && d->loss != JPCB_LOSS_CE) return fail(JPCB_EINVAL, "`mse` and `ce` are the only valid losses.");
  if (d->dtype != JPCB_DTYPE_F32 && d->dtype != JPCB_DTYPE_BF16 && d->dtype != JPCB_DTYPE_F32X3) return fail(JPCB_EINVAL, "bad dtype %d", d->dtype);
  if (d->solver != JPCB_SOLVER_EULER && d->solver != JPCB_SOLVER_HEUN) return fail(JPCB_EINVAL, "bad solver %d", d->solver);
  if (d->n_steps < 0) return fail(JPCB_EINVAL, "n_steps=%d", d->n_steps);
  if (d->dtype != JPCB_DTYPE_F32) {
    if (d->batch_local % 8) return fail(JPCB_ENOTIMPL, "the tensor-core paths (bf16, f32x3) need batch_local %% 8 == 0 (got %d)", d->batch_local);
    for (int l = 0; l <= d->n_layers; ++l)
      if (d->dims[l] % 8) return fail(JPCB_ENOTIMPL, "the tensor-core paths (bf16, f32x3) need every dim %% 8 == 0 (dims[%d]=%d)", l, d->dims[l]);
  }
  return JPCB_OK;
}

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
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

constexpr int RED_FLOATS = JPCB_MAX_LAYERS + 64;  // [0,L): layer energy sums; then misc accumulators
constexpr int RED_LOSS = JPCB_MAX_LAYERS + 0;
constexpr int RED_SUMSQ = JPCB_MAX_LAYERS + 1;

// Operand-copy geometry of the tensor-core paths.  bf16: plain copies [rows, d].  f32x3 (fp32 parity): 3-way split copies
// [rows, 3 P(d)] = [hi | mid | lo], P(d) = d rounded up to the 64-wide k-block with zero padding (elementwise.cuh).
inline int x3_part(const jpcb_pc_desc* d, int dd) { return d->dtype == JPCB_DTYPE_F32X3 ? (dd + 63) / 64 * 64 : dd; }
inline size_t x3_cw(const jpcb_pc_desc* d, int dd) { return d->dtype == JPCB_DTYPE_F32X3 ? 3 * (size_t)x3_part(d, dd) : (size_t)dd; }
// any layer of the form post_act(W act(u) + b) with a non-identity post_act (Form B)
inline bool any_post(const jpcb_pc_desc* d) {
  for (int l = 0; l < d->n_layers; ++l) if (d->post_act[l] != JPCB_ACT_LINEAR) return true;
  return false;
}

struct Ws {
  float* red = nullptr;
  std::vector<float*> err;   // [L]  err[l]: [Bl, d_{l+1}] fp32 (F32 path; tensor-core paths: f32x3, or any Form B layer)
  std::vector<float*> eff;   // [L]  Form B layers on the fp32 / f32x3 paths: post'(pre_l) . err[l] (GEMM operand, bias gradient)
  float* P1 = nullptr;       // hoisted layer-0 prediction [Bl, d_1]
  float* logits = nullptr;   // cross-entropy output: the output layer's prediction [Bl, d_L]
  std::vector<float*> Zt;    // [L-1] Heun stage state of A[l]
  std::vector<float*> G1;    // [L-1] Heun first-stage gradient of A[l]
  float* Zx = nullptr;       // free_input + Heun: stage state / first-stage gradient of z_0
  float* G1x = nullptr;
  // fp32 operand copies (fp32 path): act_l(u_l) [Bl, d_l] (+ Heun stage twin) and W_l^T [d_l, d_{l+1}], so that
  // every inference GEMM reads two K-contiguous fp32 operands with no activation left to apply on load
  std::vector<float*> Hf, Hft, WTf;
  // bf16 operand copies (tensor-core path)
  std::vector<bf16*> Wb, WTb;  // [L]   W_l [d_{l+1}, d_l] and its transpose [d_l, d_{l+1}]
  std::vector<bf16*> Hb;       // [L]   act_l(u_l)  [Bl, d_l]
  std::vector<bf16*> Htb;      // [L]   same for the Heun stage state
  std::vector<bf16*> Eb;       // [L]   err[l] [Bl, d_{l+1}]
  float* fix_part = nullptr;   // partial sums of the fixed-order reductions (jpcb_set_reproducible): (L + 2) jobs x FIX_BLOCKS
  int* sched = nullptr;        // completion counters of the whole-solve launch: 2 (L-1) tasks x row blocks of 256
  size_t sched_ints = 0;
  size_t bytes = 0;
};

void carve(const jpcb_pc_desc* d, void* base, Ws* w) {
  const int L = d->n_layers;
  const size_t Bl = (size_t)d->batch_local;
  Bump b{reinterpret_cast<char*>(base)};
  w->red = b.take<float>(RED_FLOATS);
  w->fix_part = b.take<float>((size_t)(L + 2) * FIX_BLOCKS);
  w->err.resize(L);
  w->Zt.assign(L, nullptr);
  w->G1.assign(L, nullptr);
  w->P1 = b.take<float>(Bl * d->dims[1]);
  if (d->loss == JPCB_LOSS_CE) w->logits = b.take<float>(Bl * d->dims[L]);
  const bool tc = d->dtype != JPCB_DTYPE_F32, x3 = d->dtype == JPCB_DTYPE_F32X3, heun = d->solver == JPCB_SOLVER_HEUN;
  // (the split path keeps fp32 errors too: e_l / skip sources of the gradient epilogue, bias gradients; so does any
  //  network with a Form B layer, whose bf16 error copies hold post'(pre) . e instead of e)
  const bool post = any_post(d);
  for (int l = 0; l < L; ++l) w->err[l] = (tc && !x3 && !post) ? nullptr : b.take<float>(Bl * d->dims[l + 1]);
  w->eff.assign(L, nullptr);
  if (!tc || x3)
    for (int l = 0; l < L; ++l)
      if (d->post_act[l] != JPCB_ACT_LINEAR) w->eff[l] = b.take<float>(Bl * d->dims[l + 1]);
  const bool fr = d->free_input != 0;
  if (heun) {
    for (int l = 0; l < L - 1; ++l) {
      w->Zt[l] = b.take<float>(Bl * d->dims[l + 1]);
      w->G1[l] = b.take<float>(Bl * d->dims[l + 1]);
    }
    if (fr) {
      w->Zx = b.take<float>(Bl * d->dims[0]);
      w->G1x = b.take<float>(Bl * d->dims[0]);
    }
  }
  w->Hf.assign(L, nullptr); w->Hft.assign(L, nullptr); w->WTf.assign(L, nullptr);
  if (!tc) {
    for (int l = 0; l < L; ++l) {
      if (l >= 1 || d->act[0] != JPCB_ACT_LINEAR) w->Hf[l] = b.take<float>(Bl * d->dims[l]);
      if (heun && (l >= 1 || (fr && w->Hf[0]))) w->Hft[l] = b.take<float>(Bl * d->dims[l]);
      if (l >= 1 || fr) w->WTf[l] = b.take<float>((size_t)d->dims[l] * d->dims[l + 1]);
    }
  }
  w->Wb.assign(L, nullptr); w->WTb.assign(L, nullptr); w->Hb.assign(L, nullptr); w->Htb.assign(L, nullptr);
  w->Eb.assign(L, nullptr);
  if (tc) {
    for (int l = 0; l < L; ++l) {
      w->Wb[l] = b.take<bf16>((size_t)d->dims[l + 1] * x3_cw(d, d->dims[l]));
      if (l >= 1 || fr) w->WTb[l] = b.take<bf16>((size_t)d->dims[l] * x3_cw(d, d->dims[l + 1]));
      w->Hb[l] = b.take<bf16>(Bl * x3_cw(d, d->dims[l]));
      if (heun && (l >= 1 || fr)) w->Htb[l] = b.take<bf16>(Bl * x3_cw(d, d->dims[l]));
      w->Eb[l] = b.take<bf16>(Bl * x3_cw(d, d->dims[l + 1]));
    }
    if (!x3) {
      w->sched_ints = (size_t)2 * L * ((Bl + 255) / 256);
      w->sched = b.take<int>(w->sched_ints);
    }
  }
  w->bytes = ((b.off + 1023) & ~(size_t)1023) + 1024;
}

// ---------------------------------------------------------------------------------------------
// CUDA-core launches (the tensor-core ones live in tc_gemm.cu)
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int BK, int RM, int RN>
int launch_simt_cfg(const std::vector<GTask>& tasks, cudaStream_t st) {
  size_t i = 0;
  while (i < tasks.size()) {
    SimtGroup g;
    memset(&g, 0, sizeof(int) * 2);
    int n = 0, tiles = 0;
    for (; i < tasks.size() && n < SIMT_MAX_TASKS; ++i, ++n) {
      const GTask& t = tasks[i];
      SimtTask& s = g.t[n];
      s.A = Opnd{t.a_p, t.a_smn, t.a_sk, t.a_act, 0};
      s.B = Opnd{t.b_p, t.b_smn, t.b_sk, t.b_act, 0};
      s.K = t.K;
      s.tile_begin = tiles;
      s.tiles_n = (t.e.N + BN - 1) / BN;
      s.pad_ = 0;
      s.e = t.e;
      tiles += s.tiles_n * ((t.e.M + BM - 1) / BM);
    }
    g.ntasks = n;
    g.total_tiles = tiles;
    if (tiles == 0) continue;
    constexpr int NT = (BM / (4 * RM)) * (BN / (4 * RN));
    simt_grouped_gemm<BM, BN, BK, RM, RN><<<tiles, NT, 0, st>>>(g); count_launch();
    CK(cudaGetLastError());
  }
  return JPCB_OK;
}

// narrow-layer path: one warp per MT x 4 output patch (warp_gemm.cuh); needs K-contiguous, activation-free operands
template <int MT>
int launch_warp_mt(const std::vector<GTask>& tasks, cudaStream_t st) {
  size_t i = 0;
  while (i < tasks.size()) {
    SimtGroup g;
    memset(&g, 0, sizeof(int) * 2);
    int n = 0, tiles = 0;
    for (; i < tasks.size() && n < SIMT_MAX_TASKS; ++i, ++n) {
      const GTask& t = tasks[i];
      SimtTask& s = g.t[n];
      s.A = Opnd{t.a_p, t.a_smn, 1, ACT_LINEAR, 0};
      s.B = Opnd{t.b_p, t.b_smn, 1, ACT_LINEAR, 0};
      s.K = t.K;
      s.tile_begin = tiles;
      s.tiles_n = (t.e.N + WG_NT - 1) / WG_NT;
      s.pad_ = 0;
      s.e = t.e;
      tiles += s.tiles_n * ((t.e.M + MT - 1) / MT);
    }
    g.ntasks = n;
    g.total_tiles = tiles;
    if (tiles == 0) continue;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((tiles + WG_WARPS - 1) / WG_WARPS);
    cfg.blockDim = dim3(WG_WARPS * 32);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // the kernel starts with griddepcontrol.wait
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, warp_grouped_gemm<MT>, g));
    count_launch();
  }
  return JPCB_OK;
}

// the 8 x 16-per-warp variant (warp8_grouped_gemm)
int launch_warp8(const std::vector<GTask>& tasks, cudaStream_t st) {
  size_t i = 0;
  while (i < tasks.size()) {
    SimtGroup g;
    memset(&g, 0, sizeof(int) * 2);
    int n = 0, tiles = 0;
    for (; i < tasks.size() && n < SIMT_MAX_TASKS; ++i, ++n) {
      const GTask& t = tasks[i];
      SimtTask& s = g.t[n];
      s.A = Opnd{t.a_p, t.a_smn, 1, ACT_LINEAR, 0};
      s.B = Opnd{t.b_p, t.b_smn, 1, ACT_LINEAR, 0};
      s.K = t.K;
      s.tile_begin = tiles;
      s.tiles_n = (t.e.N + WG8_NT - 1) / WG8_NT;
      s.pad_ = 0;
      s.e = t.e;
      tiles += s.tiles_n * ((t.e.M + WG8_MT - 1) / WG8_MT);
    }
    g.ntasks = n;
    g.total_tiles = tiles;
    if (tiles == 0) continue;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((tiles + WG_WARPS - 1) / WG_WARPS);
    cfg.blockDim = dim3(WG_WARPS * 32);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, warp8_grouped_gemm, g));
    count_launch();
  }
  return JPCB_OK;
}

int launch_warp(const std::vector<GTask>& tasks, cudaStream_t st) {
  long long w8 = 0, w16 = 0;
  bool small_idx = true;
  int kmax = 0;
  for (const GTask& t : tasks) {
    kmax = t.K > kmax ? t.K : kmax;
    w8 += (long long)((t.e.M + 7) / 8) * ((t.e.N + WG_NT - 1) / WG_NT);
    w16 += (long long)((t.e.M + WG8_MT - 1) / WG8_MT) * ((t.e.N + WG8_NT - 1) / WG8_NT);
    small_idx = small_idx && (long long)t.e.M * t.a_smn < (1LL << 31) && (long long)t.e.N * t.b_smn < (1LL << 31);
  }
  // enough 8 x 16 warp tiles for several warps per SM: the wide variant amortises the per-warp fixed costs 4x
  static const int w16_q = [] { const char* v = getenv("JPCB_WARP8_Q"); return v ? atoi(v) : 6; }();
  // ... as long as K is short: with only 8 lanes on K a long reduction is a long chain of dependent L2 round trips
  static const int w16_k = [] { const char* v = getenv("JPCB_WARP8_KMAX"); return v ? atoi(v) : 512; }();
  if (small_idx && w16_q != 0 && kmax <= w16_k && w16 >= (long long)w16_q * num_sms())
    return launch_warp8(tasks, st);  // (negative JPCB_WARP8_Q: always, for tests)
  // 8-row patches move less operand data per multiply-add; use them while they still give every SM >= 16 warps
  return w8 >= 16LL * num_sms() ? launch_warp_mt<8>(tasks, st) : launch_warp_mt<4>(tasks, st);
}

}  // namespace

int jpcb::launch_simt(const std::vector<GTask>& tasks, cudaStream_t st) {
  if (tasks.empty()) return JPCB_OK;
  // pick the kernel / tile shape from the amount of parallelism the phase offers
  long long t128 = 0, t64 = 0;
  bool warp_ok = !(tc_flags() & 32);
  for (const GTask& t : tasks) {
    t128 += (long long)((t.e.M + 127) / 128) * ((t.e.N + 127) / 128);
    t64 += (long long)((t.e.M + 63) / 64) * ((t.e.N + 63) / 64);
    warp_ok = warp_ok && (t.K == 0 || (t.a_sk == 1 && t.b_sk == 1 && t.a_act == ACT_LINEAR && t.b_act == ACT_LINEAR));
  }
  const int sms = num_sms();
  // fewer 64 x 64 tiles than one per SM: a tiled GEMM cannot fill the machine, the warp-level path can
  static const int warp_q = [] { const char* v = getenv("JPCB_WARP_Q"); return v ? atoi(v) : 4; }();  // threshold in quarter-SMs
  if (warp_ok && 4 * t64 < (long long)warp_q * sms) return launch_warp(tasks, st);
  if (t128 >= 2LL * sms) return launch_simt_cfg<128, 128, 16, 2, 2>(tasks, st);
  if (t64 >= sms) return launch_simt_cfg<64, 64, 16, 1, 1>(tasks, st);
  return launch_simt_cfg<32, 32, 16, 1, 1>(tasks, st);  // (8x8 threads) latency-bound small layers
}

namespace {

int ew_grid(size_t work_items, int per_block) {
  size_t blocks = (work_items + per_block - 1) / per_block;
  const size_t cap = (size_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
// one invocation: descriptor + caller buffers + carved workspace
// ---------------------------------------------------------------------------------------------
struct Run {
  const jpcb_pc_desc* d;
  const float* const* W;
  const float* const* b;
  const float* x;        // the input -- or, with free_input, z_0 of the current state
  const float* y;
  Ws ws;
  cudaStream_t st;
  int L, Bl;
  float invB;
  bool tc;
  bool x3 = false;                // JPCB_DTYPE_F32X3: split operand copies, six products per GEMM
  bool fr = false;                // desc->free_input: z_0 is a free activity (unsupervised PC)
  bool post = false;              // some layer is of the form post_act(W act(u) + b) (Form B)
  bool hoist = true;              // the layer-0 prediction is constant during inference and kept as P1 (needs an input and a
                                  // Form A first layer: a Form B one also needs post'(pre_0) for its weight gradient)
  const float* x_stage = nullptr; // free_input: z_0 of the Heun stage state (else the input)
  float* x_out = nullptr;         // free_input: where the next grads() call writes z_0's gradient / new value

  int init(const jpcb_pc_desc* desc, const float* const* W_, const float* const* b_, const float* x_, const float* y_, void* wsp,
           size_t ws_bytes, void* stream, int min_layers = 2) {
    RET(validate(desc, min_layers));
    d = desc; W = W_; b = b_; x = x_; y = y_;
    st = reinterpret_cast<cudaStream_t>(stream);
    L = d->n_layers; Bl = d->batch_local; invB = 1.f / (float)d->batch_global;
    tc = d->dtype != JPCB_DTYPE_F32; x3 = d->dtype == JPCB_DTYPE_F32X3;
    if (!W || !x) return fail(JPCB_EINVAL, "null weight list or input");
    if (d->has_bias && !b) return fail(JPCB_EINVAL, "has_bias set but no bias list");
    carve(d, wsp, &ws);
    if (!wsp || ws_bytes < ws.bytes) return fail(JPCB_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", ws.bytes, ws_bytes);
    fr = d->free_input != 0;
    post = any_post(d);
    hoist = !fr && d->post_act[0] == JPCB_ACT_LINEAR;
    x_stage = fr && ws.Zx ? ws.Zx : x;
    return JPCB_OK;
  }
  const float* bias(int l) const { return d->has_bias ? b[l] : nullptr; }
  // split-path geometry of a tensor-core task whose concatenated dimension is `ka` (A) / `kb` (B) wide, and of a bf16 output copy
  void split_geom(GTask& t, int ka, int kb) const {
    if (!x3) return;
    t.split = 6; t.a_ld = (int)x3_cw(d, ka); t.b_ld = (int)x3_cw(d, kb); t.a_part = x3_part(d, ka); t.b_part = x3_part(d, kb);
  }
  void set_ob(GTask& t, bf16* p) const { t.e.Ob = p; if (x3 && p) { t.e.ob_ld = (int)x3_cw(d, t.e.N); t.e.ob_part = x3_part(d, t.e.N); } }
  int split3(const float* src, bf16* dst, int R, int C, int act) {
    const int P = x3_part(d, C);
    split3_kernel<<<ew_grid((size_t)R * P, 256), 256, 0, st>>>(src, dst, R, C, P, act); count_launch();
    return JPCB_OK;
  }
  // bf16 operand copy of a [Bl, C] fp32 tensor: plain cast or 3-way split
  int operand_copy(const float* src, bf16* dst, int C, int act) {
    if (x3) return split3(src, dst, Bl, C, act);
    const size_t n = (size_t)Bl * C;
    cast_act_bf16_kernel<<<ew_grid(n / 4 + 1, 256), 256, 0, st>>>(src, dst, n, act); count_launch();
    return JPCB_OK;
  }
  float eskip(int l) const { return (d->skip[l] && l >= 1 && l <= L - 2) ? 1.f : 0.f; }  // energy-side skip (_energies.py:111-117)
  int launch(const std::vector<GTask>& t) { return tc ? launch_tc(t, st) : launch_simt(t, st); }

  int zero_red() { CK(cudaMemsetAsync(ws.red, 0, RED_FLOATS * sizeof(float), st)); return JPCB_OK; }

  // fixed-order sums into ws.red (jpcb_set_reproducible): the prediction-error energy of layer l from its stored errors,
  // out = lam d  =>  lam/2 sum d^2 = sum out^2 / (2 lam); the mse loss of the init from the network output and the target
  std::vector<FixJob> fix;
  void fix_energy(int l, float escale) {
    const bool b16 = ws.err[l] == nullptr;   // (plain bf16 path: only the bf16 error copies exist)
    fix.push_back(FixJob{b16 ? (const void*)ws.Eb[l] : (const void*)ws.err[l], nullptr, (unsigned long long)Bl * d->dims[l + 1],
                         escale != 0.f ? 0.5f / escale : 0.f, b16 ? 1 : 0, l, 0});
  }
  int fix_flush() {
    if (fix.empty()) return JPCB_OK;
    if (fix.size() > (size_t)(L + 2) || fix.size() > FIX_MAX_JOBS) return fail(JPCB_EINVAL, "internal: %zu fixed-order sums", fix.size());
    static thread_local FixJobs jobs;
    jobs.n = (int)fix.size(); jobs.pad_ = 0;
    for (size_t i = 0; i < fix.size(); ++i) jobs.j[i] = fix[i];
    fix_partial_kernel<<<dim3(FIX_BLOCKS, jobs.n), 256, 0, st>>>(jobs, ws.fix_part); count_launch();
    fix_final_kernel<<<1, 256, 0, st>>>(jobs, ws.fix_part, ws.red); count_launch();
    CK(cudaGetLastError());
    fix.clear();
    return JPCB_OK;
  }

  // ---- bf16 operand preparation ----
  int prep_weights() {
    if (!tc) {  // W_l^T copies for the transposed GEMMs of the gradient phases: every layer in one grouped launch
      static thread_local TrJobs jobs;
      jobs.n = 0; jobs.total = 0;
      auto flush = [&]() {
        if (jobs.n == 0) return;
        transpose_f32_group_kernel<<<ew_grid(jobs.total, 1), dim3(32, 8), 0, st>>>(jobs); count_launch();
        jobs.n = 0; jobs.total = 0;
      };
      for (int l = fr ? 0 : 1; l < L; ++l) {
        if (jobs.n == TR_MAX_JOBS) flush();
        TrJob& jb = jobs.j[jobs.n++];
        jb.src = W[l]; jb.dst = ws.WTf[l]; jb.R = d->dims[l + 1]; jb.C = d->dims[l];
        jb.tile_begin = jobs.total; jb.tiles_c = (jb.C + 31) / 32;
        jobs.total += jb.tiles_c * ((jb.R + 31) / 32);
      }
      flush();
      CK(cudaGetLastError());
      return JPCB_OK;
    }
    if (x3) {
      for (int l = 0; l < L; ++l) {
        const int R = d->dims[l + 1], C = d->dims[l], PR = x3_part(d, R);
        split3(W[l], ws.Wb[l], R, C, ACT_LINEAR);
        if (l >= 1 || fr) {  // W_l [R, C] -> split3(W_l^T) [C, 3 P(R)]
          transpose_split3_kernel<<<ew_grid(((PR + 31) / 32) * ((C + 31) / 32), 1), dim3(32, 8), 0, st>>>(W[l], ws.WTb[l], R, C, PR); count_launch();
        }
        // copies the epilogues write (valid columns only): their padding columns must read as zero
        CK(cudaMemsetAsync(ws.Eb[l], 0, (size_t)Bl * x3_cw(d, R) * sizeof(bf16), st));
        CK(cudaMemsetAsync(ws.Hb[l], 0, (size_t)Bl * x3_cw(d, C) * sizeof(bf16), st));
        if (ws.Htb[l]) CK(cudaMemsetAsync(ws.Htb[l], 0, (size_t)Bl * x3_cw(d, C) * sizeof(bf16), st));
      }
      CK(cudaGetLastError());
      return JPCB_OK;
    }
    // plain bf16 path: every layer's Wb (and W^T copy) in one grouped launch, each weight read once
    for (int l0 = 0; l0 < L; l0 += PREP_MAX_JOBS) {
      PrepJobs jobs;
      jobs.n = 0; jobs.total = 0;
      for (int l = l0; l < L && jobs.n < PREP_MAX_JOBS; ++l) {
        PrepJob& jb = jobs.j[jobs.n++];
        jb.W = W[l]; jb.Wb = ws.Wb[l]; jb.WTb = (l >= 1 || fr) ? ws.WTb[l] : nullptr;
        jb.R = d->dims[l + 1]; jb.C = d->dims[l];
        jb.tile_begin = jobs.total; jb.tiles_c = (jb.C + 63) / 64;
        jobs.total += jb.tiles_c * ((jb.R + 63) / 64);
      }
      prep_weights_kernel<<<ew_grid(jobs.total, 1), 256, 0, st>>>(jobs); count_launch();
    }
    CK(cudaGetLastError());
    return JPCB_OK;
  }
  // operand copies of the layer inputs on state S: (Hb | Hf)[0] = act_0(x); when `S` is given also
  // (Hb | Hf)[l] = act_l(S[l-1]) for l = 1..L-1  (bf16 on the tensor-core path, fp32 otherwise)
  int prep_inputs(const float* const* S) {
    if (!tc) {
      if (ws.Hf[0]) { const size_t n0 = (size_t)Bl * d->dims[0]; act_f32_kernel<<<ew_grid(n0, 256), 256, 0, st>>>(x, ws.Hf[0], n0, d->act[0]); count_launch(); }
      if (S)
        for (int l = 1; l < L; ++l) {
          const size_t n = (size_t)Bl * d->dims[l];
          act_f32_kernel<<<ew_grid(n, 256), 256, 0, st>>>(S[l - 1], ws.Hf[l], n, d->act[l]); count_launch();
        }
      CK(cudaGetLastError());
      return JPCB_OK;
    }
    operand_copy(x, ws.Hb[0], d->dims[0], d->act[0]);
    if (S)
      for (int l = 1; l < L; ++l) operand_copy(S[l - 1], ws.Hb[l], d->dims[l], d->act[l]);
    CK(cudaGetLastError());
    return JPCB_OK;
  }

  // ---- task builders ----
  // forward GEMM of layer l on state S (u_l = x or S[l-1]); H = bf16 operand list for that state
  // `stage`: the operand copies of the Heun stage state (Htb / Hft) instead of the current state's
  void fwd_operands(GTask& t, int l, const float* const* S, bool stage) const {
    const std::vector<bf16*>& H = stage ? ws.Htb : ws.Hb;
    const std::vector<float*>& Hf = stage ? ws.Hft : ws.Hf;
    t.e.M = Bl; t.e.N = d->dims[l + 1]; t.K = d->dims[l];
    (void)S;
    const bool st0 = stage && fr;  // layer 0 reads the stage state's z_0 only when z_0 is free
    t.a_p = l == 0 ? (st0 ? (ws.Hft[0] ? ws.Hft[0] : x_stage) : (ws.Hf[0] ? ws.Hf[0] : x)) : Hf[l];
    t.a_smn = d->dims[l]; t.a_sk = 1; t.a_act = ACT_LINEAR;
    t.b_p = W[l]; t.b_smn = d->dims[l]; t.b_sk = 1;
    t.a_b = l == 0 ? (st0 ? ws.Htb[0] : ws.Hb[0]) : H[l]; t.b_b = ws.Wb[l];
    split_geom(t, d->dims[l], d->dims[l]);
    t.e.alpha = d->scalings[l];
    t.e.bias = bias(l);
    t.e.pre_act = d->post_act[l];
  }

  // The error wavefront (infer_loop): from the feed-forward init every hidden error is exactly zero and Euler steps move it
  // one layer per step.  (a skip at layer 0 enters the feed-forward init but not the energy, _init.py:69-78 vs
  // _energies.py:111-126: then err[0] = x != 0 at the init and there is no wavefront to exploit)
  bool wavefront_ok() const {
    return hoist && !post && d->solver == JPCB_SOLVER_EULER && d->activity_decay == 0.f && L > 2 && !d->skip[0] && !(tc_flags() & 8);
  }

  // deep narrow networks on the fp32 path: the whole feed-forward chain as ONE launch in which a CTA owns a few batch rows
  // and walks all layers for them (chain_ffwd.cuh).  Worth it when the chain is long and every CTA can afford to stream
  // every weight from L2: L launches of a few microseconds each against (sum of weights) / (~32 B/clk) per CTA.
  bool chain_ffwd_pays() const {
    const char* v = getenv("JPCB_CHAIN_FFWD");  // 0: never, 2: whenever legal (read per call: tests compare both paths)
    const int mode = v ? atoi(v) : 1;
    if (tc || mode == 0 || L > CF_MAX_LAYERS) return false;
    double wbytes = 0.0;
    for (int l = 0; l < L; ++l) {
      if (d->dims[l] > 1024 || d->dims[l + 1] > 1024) return false;   // (shared memory: 4 x R rows of the widest layer)
      wbytes += 4.0 * d->dims[l] * d->dims[l + 1];
    }
    if (mode == 2) return true;
    return L >= 6 && wbytes / 60e3 < 3.0 * L;  // microseconds: ~60 GB/s-per-SM streaming vs ~3 us per dependent launch
  }
  int chain_ffwd(float* const* A_out, float* loss_acc, bool keep_p1) {
    static thread_local CfArgs a;
    a.L = L; a.B = Bl; a.x = x;
    const bool ce = d->loss == JPCB_LOSS_CE;
    const bool fix_loss = loss_acc && y && !ce && reproducible() && loss_acc == ws.red + RED_LOSS;
    const bool want_loss = loss_acc && y && !ce && !fix_loss;
    a.Y = want_loss ? y : nullptr; a.loss = want_loss ? loss_acc : nullptr;
    int kmax = 0;
    for (int l = 0; l < L; ++l) {
      CfLayer& c = a.l[l];
      c.W = W[l]; c.bias = bias(l); c.O = A_out[l];
      c.O2 = (l == 0 && keep_p1) ? ws.P1 : nullptr;
      c.Oh = l + 1 < L ? ws.Hf[l + 1] : nullptr;
      c.K = d->dims[l]; c.N = d->dims[l + 1];
      c.act_in = d->act[l]; c.act_post = d->post_act[l]; c.act_next = l + 1 < L ? d->act[l + 1] : ACT_LINEAR;
      c.alpha = d->scalings[l]; c.skip = d->skip[l] ? 1.f : 0.f;  // the init honours any non-None skip (_init.py:69-78)
      kmax = c.K > kmax ? c.K : kmax;
    }
    for (int l = 0; l < L; ++l) kmax = d->dims[l + 1] > kmax ? d->dims[l + 1] : kmax;
    const int kpad = (kmax + 7) / 4 * 4;
    constexpr int R = 4;
    const size_t smem = (size_t)4 * R * kpad * sizeof(float);  // raw + act, double-buffered
    {  // (per-device function attribute, as for the tensor-core kernels)
      static std::atomic<unsigned long long> done{0};
      int dev = 0;
      CK(cudaGetDevice(&dev));
      const unsigned long long bit = dev < 64 ? (1ull << dev) : 0ull;
      if (!bit || !(done.load(std::memory_order_relaxed) & bit)) {
        CK(cudaFuncSetAttribute(chain_ffwd_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * R * 1032 * (int)sizeof(float)));
        if (bit) done.fetch_or(bit, std::memory_order_relaxed);
      }
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((Bl + R - 1) / R);
    cfg.blockDim = dim3(CF_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // the kernel starts with griddepcontrol.wait
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, chain_ffwd_kernel<R>, a, kpad));
    count_launch();
    if (fix_loss) {
      fix.push_back(FixJob{A_out[L - 1], y, (unsigned long long)Bl * d->dims[L], 0.5f, 0, RED_LOSS, 0});
      RET(fix_flush());
    }
    if (loss_acc && y && ce) RET(ce_rows(A_out[L - 1], 1.f, nullptr, nullptr, loss_acc));
    return JPCB_OK;
  }

  // init_activities_with_ffwd (jpc/_core/_init.py:68-80); layer 0 also leaves the hoisted P1
  // `chain_ok`: the caller does not rely on a later error pass reproducing these activities BIT for bit (the chain kernel
  // sums in another order than the error kernels: "z_l - prediction" of an untouched layer is then ~1e-9 instead of 0)
  int ffwd(float* const* A_out, float* loss_acc, bool keep_p1, bool chain_ok) {
    if (chain_ok && chain_ffwd_pays()) return chain_ffwd(A_out, loss_acc, keep_p1);
    for (int l = 0; l < L; ++l) {
      GTask t;
      fwd_operands(t, l, A_out, false);
      t.e.mode = EPI_FFWD;
      t.e.skip = d->skip[l] ? 1.f : 0.f;  // the init honours any non-None skip (_init.py:69-78)
      t.e.U = l == 0 ? x : A_out[l - 1];
      t.e.O = A_out[l];
      if (l == 0 && keep_p1) t.e.O2 = ws.P1;
      if (l + 1 < L) { t.e.act = d->act[l + 1]; if (tc) set_ob(t, ws.Hb[l + 1]); else t.e.Oh = ws.Hf[l + 1]; }
      const bool ce = d->loss == JPCB_LOSS_CE;
      const bool fix_loss = l == L - 1 && loss_acc && y && !ce && reproducible() && loss_acc == ws.red + RED_LOSS;
      if (l == L - 1 && loss_acc && y && !ce && !fix_loss) { t.e.red = loss_acc; t.e.Y = y; }
      RET(launch({t}));
      if (fix_loss) {
        fix.push_back(FixJob{A_out[L - 1], y, (unsigned long long)Bl * d->dims[L], 0.5f, 0, RED_LOSS, 0});
        RET(fix_flush());
      }
      // cross_entropy_loss(logits = network output, y)  (jpc/_utils.py:133-136)
      if (l == L - 1 && loss_acc && y && ce) RET(ce_rows(A_out[L - 1], 1.f, nullptr, nullptr, loss_acc));
    }
    return JPCB_OK;
  }

  // hoisted layer-0 prediction P1 = a_0 (W_0 act_0(x) + b_0)  (constant during inference)
  int predict_p1() {
    GTask t;
    fwd_operands(t, 0, nullptr, false);
    t.e.mode = EPI_FFWD;
    t.e.O = ws.P1;
    return launch({t});
  }

  // prediction errors err[l] for l in [lo, L) on state S; optional per-layer energy sums
  // one warp per row: cross-entropy error / energy of the output layer from its logits
  int ce_rows(const float* lg, float lam, float* err, bf16* err_b, float* red) {
    const int wpb = 8, blocks = (Bl + wpb - 1) / wpb;
    ce_rows_kernel<<<blocks < num_sms() * 4 ? blocks : num_sms() * 4, wpb * 32, 0, st>>>(lg, y, Bl, d->dims[L], lam, err, x3 ? nullptr : err_b, red);
    count_launch();
    if (x3 && err_b && err) split3(err, err_b, Bl, d->dims[L], ACT_LINEAR);  // (the split copy is made from the fp32 error)
    CK(cudaGetLastError());
    return JPCB_OK;
  }

  // `Tl` / `esc_all` (layer-wise regression, jpcb_hpc_param_grads): per-layer targets instead of the next
  // activity, and one error weight for every layer instead of lambda on the output layer only
  // `collect`: hand the tasks back instead of launching them (the whole-solve launch takes them as its schedule)
  int errors(int lo, const float* const* S, bool stage, bool energy, const float* const* Tl = nullptr, float esc_all = 0.f,
             std::vector<GTask>* collect = nullptr) {
    std::vector<GTask> ts;
    const bool ce = d->loss == JPCB_LOSS_CE;
    for (int l = lo; l < L; ++l) {
      GTask t;
      fwd_operands(t, l, S, stage);
      if (ce && l == L - 1) {  // logits only; the softmax error follows in ce_rows (needs whole rows)
        t.e.mode = EPI_FFWD;
        t.e.O = ws.logits;
        ts.push_back(t);
        continue;
      }
      t.e.mode = EPI_ERR;
      t.e.skip = eskip(l);
      t.e.U = l == 0 ? x : S[l - 1];
      t.e.T = Tl ? Tl[l] : (l == L - 1 ? y : S[l]);
      t.e.escale = esc_all != 0.f ? esc_all : (l == L - 1 ? d->output_energy_scaling : 1.f);
      t.e.O = ws.err[l];
      set_ob(t, ws.Eb[l]);
      t.e.Oh = ws.eff[l];  // (Form B, fp32 / f32x3 paths: post'(pre) . e)
      if (energy && reproducible() && !collect) fix_energy(l, t.e.escale);   // summed after the launch, in a fixed order
      else if (energy) t.e.red = ws.red + l;
      ts.push_back(t);
    }
    if (collect) { *collect = ts; return JPCB_OK; }
    RET(launch(ts));
    RET(fix_flush());
    if (ce) RET(ce_rows(ws.logits, d->output_energy_scaling, ws.err[L - 1], ws.Eb[L - 1], energy ? ws.red + (L - 1) : nullptr));
    return JPCB_OK;
  }
  // err[0] = S[0] - P1 without a GEMM (K = 0 task on the CUDA-core kernel), optional energy
  int error0_from_p1(const float* const* S, bool energy) {
    const size_t n0 = (size_t)Bl * d->dims[1];
    if (n0 % 4 == 0 && (reinterpret_cast<uintptr_t>(S[0]) & 15) == 0) {  // plain streaming pass (the usual case)
      const bool fixed = energy && reproducible() && (ws.err[0] || !x3);
      err0_kernel<<<ew_grid(n0 / 4, 256), 256, 0, st>>>(S[0], ws.P1, ws.err[0], ws.Eb[0], n0 / 4, energy && !fixed ? ws.red + 0 : nullptr, d->dims[1],
                                                         x3 ? (int)x3_cw(d, d->dims[1]) : 0, x3 ? x3_part(d, d->dims[1]) : 0);
      count_launch();
      CK(cudaGetLastError());
      if (fixed) { fix_energy(0, 1.f); RET(fix_flush()); }
      return JPCB_OK;
    }
    GTask t;
    t.e.mode = EPI_ERR;
    t.e.M = Bl; t.e.N = d->dims[1]; t.K = 0;
    t.e.alpha = 0.f; t.e.skip = 1.f; t.e.U = ws.P1; t.e.T = S[0];
    t.e.O = ws.err[0]; set_ob(t, ws.Eb[0]);
    const bool fixed = energy && reproducible();
    if (energy && !fixed) t.e.red = ws.red + 0;
    t.a_p = S[0]; t.b_p = S[0];  // never dereferenced (K == 0)
    RET(launch_simt({t}, st));
    if (fixed) { fix_energy(0, 1.f); RET(fix_flush()); }
    return JPCB_OK;
  }

  // gradient/update of A[l-1], l = 1..L-1.  Z: state where act' is evaluated; p1_hoisted: err[0]
  // is formed on the fly as Z[0] - P1.
  // `stage_out`: the updated state is the Heun stage state (operand copies go to Htb / Hft)
  // `err_red` (GRAD_HEUN2 only): accumulates the scaled embedded-error sum of squares of an adaptive trial step
  int grads(int sub, float c, const float* const* Z, const float* const* Z0, float* const* Out, bool stage_out,
            bool p1_hoisted, int lo = 1, float* err_red = nullptr, float rtol = 0.f, float atol = 0.f,
            std::vector<GTask>* collect = nullptr) {
    std::vector<GTask> ts;
    for (int l = lo; l < L; ++l) {
      GTask t;
      t.e.mode = EPI_GRAD; t.e.sub = sub; t.e.act = d->act[l];
      t.e.M = Bl; t.e.N = d->dims[l]; t.K = d->dims[l + 1];
      t.a_p = ws.eff[l] ? ws.eff[l] : ws.err[l]; t.a_smn = d->dims[l + 1]; t.a_sk = 1;  // (Form B: post'(pre_l) . err[l] travels through W_l)
      t.b_p = ws.WTf[l]; t.b_smn = d->dims[l + 1]; t.b_sk = 1;  // B(n,k) = W_l[k][n] = WTf_l[n][k]
      t.a_b = ws.Eb[l]; t.b_b = ws.WTb[l];
      split_geom(t, d->dims[l + 1], d->dims[l + 1]);
      t.e.alpha = d->scalings[l];
      t.e.skip = eskip(l);
      // (split path: the bf16 error copies are 3-part operand copies, not readable as plain bf16 -- fp32 errors instead;
      //  Form B: the bf16 copies hold post'(pre) . e, the own-error and skip terms need the raw e)
      const bool raw32 = x3 || post;
      t.e.En = ws.err[l]; t.e.Enb = raw32 ? nullptr : ws.Eb[l];
      t.e.Z = Z[l - 1];
      if (l == 1 && p1_hoisted) t.e.P = ws.P1;
      else { t.e.El = ws.err[l - 1]; t.e.Elb = raw32 ? nullptr : ws.Eb[l - 1]; }
      t.e.c = c; t.e.ad = d->activity_decay; t.e.gscale = 1.f;
      t.e.O = Out[l - 1];
      if (sub == GRAD_HEUN1 || sub == GRAD_HEUN2) t.e.G1 = ws.G1[l - 1];
      if (sub == GRAD_HEUN2) { t.e.Z0 = Z0[l - 1]; t.e.red = err_red; t.e.tol_r = rtol; t.e.tol_a = atol; }
      if (sub != GRAD_OUT) { if (tc) set_ob(t, (stage_out ? ws.Htb : ws.Hb)[l]); else t.e.Oh = (stage_out ? ws.Hft : ws.Hf)[l]; }
      ts.push_back(t);
    }
    if (fr) {
      // unsupervised PC: z_0 is free.  It has no prediction error of its own, so its gradient is only the
      // term sent down by layer 0: dF/dz_0 = -(1/B) act_0'(z_0) . (a_0 e_1 W_0)  (+ decay).  "e_0 = 0" is
      // expressed with the hoisted-prediction source: e = Z - P with P = Z.
      if (!x_out) return fail(JPCB_EINVAL, "free_input: no destination for z_0");
      const float* zx = sub == GRAD_HEUN2 ? x_stage : x;
      GTask t;
      t.e.mode = EPI_GRAD; t.e.sub = sub; t.e.act = d->act[0];
      t.e.M = Bl; t.e.N = d->dims[0]; t.K = d->dims[1];
      t.a_p = ws.err[0]; t.a_smn = d->dims[1]; t.a_sk = 1;
      t.b_p = ws.WTf[0]; t.b_smn = d->dims[1]; t.b_sk = 1;
      t.a_b = ws.Eb[0]; t.b_b = ws.WTb[0];
      split_geom(t, d->dims[1], d->dims[1]);
      t.e.alpha = d->scalings[0];
      t.e.Z = zx; t.e.P = zx;
      t.e.c = c; t.e.ad = d->activity_decay; t.e.gscale = 1.f;
      t.e.O = x_out;
      if (sub == GRAD_HEUN1 || sub == GRAD_HEUN2) t.e.G1 = ws.G1x;
      if (sub == GRAD_HEUN2) { t.e.Z0 = x; t.e.red = err_red; t.e.tol_r = rtol; t.e.tol_a = atol; }
      if (sub != GRAD_OUT) { if (tc) set_ob(t, (stage_out ? ws.Htb : ws.Hb)[0]); else t.e.Oh = (stage_out ? ws.Hft : ws.Hf)[0]; }
      ts.push_back(t);
      x_out = nullptr;
    }
    if (collect) { *collect = ts; return JPCB_OK; }
    return launch(ts);
  }

  int scale_inplace(float* p, size_t n, float s) {
    scale_kernel<<<ew_grid(n, 256), 256, 0, st>>>(p, p, n, s); count_launch();
    CK(cudaGetLastError());
    return JPCB_OK;
  }

  // T fixed steps of Euler / Heun, in place on A (jpc/_core/_infer.py:109-132; diffrax Euler.step /
  // Heun tableau).  Requires P1 and, on the tensor-core path, Hb[l] = act_l(A[l-1]).
  //
  // `from_ffwd`: A is exactly the feed-forward init (make_pc_step).  Then every hidden prediction error is
  // exactly zero and only the output error is not, so after s Euler steps only the top s+1 activities have
  // moved (the error wavefront travels one layer per step).  Step s therefore only
  // evaluates layers l >= L-1-s; the error buffer just below the front is read as the zeros it was set to.
  // Same result as evaluating every layer (whose contributions would be exact zeros), fewer GEMMs.
  int infer_loop(float* const* A, bool from_ffwd = false) {
    const int T = d->n_steps;
    const size_t n_dummy = (size_t)Bl * d->dims[L];
    float* const xw = const_cast<float*>(x);  // free_input: z_0 is updated in place like every other activity
    // (a skip at layer 0 enters the feed-forward init but not the energy, _init.py:69-78 vs _energies.py:111-126: then
    //  err[0] = x != 0 at the init and there is no wavefront to exploit)
    const bool front = from_ffwd && wavefront_ok();
    if (front) {
      if (tc) {
        for (int l = (L - 2 - T > 0 ? L - 2 - T : 0); l < L - 1; ++l) {
          CK(cudaMemsetAsync(ws.Eb[l], 0, (size_t)Bl * x3_cw(d, d->dims[l + 1]) * sizeof(bf16), st));
          if (x3) CK(cudaMemsetAsync(ws.err[l], 0, (size_t)Bl * d->dims[l + 1] * sizeof(float), st));  // (the split path reads fp32 errors)
        }
      } else {
        CK(cudaMemsetAsync(ws.err[0], 0, (size_t)(reinterpret_cast<char*>(ws.err[L - 1]) - reinterpret_cast<char*>(ws.err[0])), st));
      }
    }
    // wide bf16 layers: all T Euler steps as ONE launch of the staged kernel (csrc/tc_staged.cuh: ERR and GRAD tiles of
    // every step in one dependency-tracked stream); anything it is not compiled for falls through to the phase loop
    if (tc && !x3 && hoist && !post && T >= 1 && d->solver == JPCB_SOLVER_EULER && d->loss == JPCB_LOSS_MSE && d->activity_decay == 0.f && ws.sched) {
      std::vector<GTask> es, gs;
      RET(errors(1, A, false, false, nullptr, 0.f, &es));
      RET(grads(GRAD_EULER, d->dt * invB, A, nullptr, A, false, true, 1, nullptr, 0.f, 0.f, &gs));
      const int rc = launch_tc_solve(es, gs, T, d->dt_last * invB, front, ws.sched, ws.sched_ints, st);
      if (rc != JPCB_SOLVE_NA) return rc;
    }
    for (int s = 0; s < T; ++s) {
      const float h = s == T - 1 ? d->dt_last : d->dt;
      const float c = h * invB;
      if (d->solver == JPCB_SOLVER_EULER) {
        const int lo = front ? (L - 1 - s > 1 ? L - 1 - s : 1) : 1;
        RET(errors(hoist ? lo : 0, A, false, false));
        x_out = fr ? xw : nullptr;
        RET(grads(GRAD_EULER, c, A, nullptr, A, false, hoist, lo));
        if (d->activity_decay != 0.f) RET(scale_inplace(A[L - 1], n_dummy, 1.f - c * d->activity_decay));
      } else {
        std::vector<float*> zt(ws.Zt.begin(), ws.Zt.begin() + (L - 1));
        zt.push_back(A[L - 1]);  // dummy slot is never read by the error tasks
        RET(errors(hoist ? 1 : 0, A, false, false));
        x_out = fr ? ws.Zx : nullptr;
        RET(grads(GRAD_HEUN1, c, A, nullptr, zt.data(), true, hoist));
        RET(errors(hoist ? 1 : 0, zt.data(), true, false));
        x_out = fr ? xw : nullptr;
        RET(grads(GRAD_HEUN2, c, zt.data(), A, A, false, hoist));
        if (d->activity_decay != 0.f) {
          const float q = c * d->activity_decay;
          RET(scale_inplace(A[L - 1], n_dummy, 1.f - q + 0.5f * q * q));
        }
      }
    }
    return JPCB_OK;
  }

  // parameter gradients from err[] (already computed on state S) -- jpc/_core/_grads.py:487-499
  int wgrads(const float* const* S, float* const* gW, float* const* gb, int l_lo = 0, int l_hi = -1) {
    if (l_hi < 0) l_hi = L;
    // tensor-core path: gW_l = -(a_l/B) Eb[l]^T Hb[l] with both operands read MN-major straight from
    // the [batch, d] bf16 copies the inference phases already keep (K = batch runs across their rows)
    std::vector<GTask> ts;
    for (int l = l_lo; l < l_hi; ++l) {
      GTask t;
      t.e.mode = EPI_SCALE;
      t.e.M = d->dims[l + 1]; t.e.N = d->dims[l]; t.K = Bl;
      t.a_p = ws.eff[l] ? ws.eff[l] : ws.err[l]; t.a_smn = 1; t.a_sk = d->dims[l + 1];  // (Form B: post'(pre_l) . err[l])
      t.b_p = l == 0 ? (ws.Hf[0] ? ws.Hf[0] : x) : ws.Hf[l]; t.b_smn = 1; t.b_sk = d->dims[l];  // act already applied
      (void)S;
      t.a_b = ws.Eb[l]; t.b_b = ws.Hb[l]; t.mn_major = tc;
      split_geom(t, d->dims[l + 1], d->dims[l]);
      t.e.alpha = -d->scalings[l] * invB;
      t.e.O = gW[l];
      ts.push_back(t);
    }
    RET(launch(ts));
    if (d->has_bias) {
      if (!gb) return fail(JPCB_EINVAL, "has_bias set but no bias-gradient list");
      for (int l = l_lo; l < l_hi; ++l) {
        const int N = d->dims[l + 1];
        const float sc = -d->scalings[l] * invB;
        if (tc && !x3) colsum_kernel<bf16><<<(N + 31) / 32, dim3(32, 8), 0, st>>>(ws.Eb[l], gb[l], Bl, N, sc);
        else colsum_kernel<float><<<(N + 31) / 32, dim3(32, 8), 0, st>>>(ws.eff[l] ? ws.eff[l] : ws.err[l], gb[l], Bl, N, sc);
        count_launch();
      }
      CK(cudaGetLastError());
    }
    return JPCB_OK;
  }

  int finalize(float* E_layers, float* F, bool with_sumsq, float* loss_out) {
    finalize_kernel<<<1, 32, 0, st>>>(ws.red, L, invB, E_layers, F, with_sumsq ? ws.red + RED_SUMSQ : nullptr, d->activity_decay,
                                      ws.red + RED_LOSS, loss_out);
    count_launch();
    CK(cudaGetLastError());
    return JPCB_OK;
  }
  int sumsq_all(const float* const* A, float* out, const float* z0 = nullptr) {
    if (fr) {  // z_0 is one of the activities
      const size_t n = (size_t)Bl * d->dims[0];
      sumsq_kernel<<<ew_grid(n, 1024), 256, 0, st>>>(z0 ? z0 : x, n, out); count_launch();
    }
    for (int l = 0; l < L; ++l) {
      const size_t n = (size_t)Bl * d->dims[l + 1];
      sumsq_kernel<<<ew_grid(n, 1024), 256, 0, st>>>(A[l], n, out); count_launch();
    }
    CK(cudaGetLastError());
    return JPCB_OK;
  }
};

// desc->free_input: the caller's activity lists have L + 1 entries [z_0, z_1 .. z_{L-1}, dummy] and there is no
// input; internally z_0 takes the input's place and the lists are seen from their second entry on.
template <typename PP>
int split_free(const jpcb_pc_desc* d, const float*& x, PP& A) {
  if (!d || !d->free_input) return JPCB_OK;
  if (x) return fail(JPCB_EINVAL, "free_input (unsupervised): `x` must be NULL, z_0 is activities[0]");
  if (!A || !A[0]) return fail(JPCB_EINVAL, "free_input (unsupervised): null activity list");
  x = A[0];
  A = A + 1;
  return JPCB_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int jpcb_version(void) { return JPCB_VERSION; }
const char* jpcb_last_error(void) { return g_err.c_str(); }
unsigned long long jpcb_launch_count(void) { return g_launches.load(); }
int jpcb_set_reproducible(int on) { g_reproducible.store(on ? 1 : 0); return JPCB_OK; }

size_t jpcb_pc_workspace_bytes(const jpcb_pc_desc* desc) {
  if (validate(desc, 1) != JPCB_OK) return 0;  // (a single layer is enough for jpcb_pc_ffwd_init)
  Ws w;
  carve(desc, nullptr, &w);
  return w.bytes;
}

static int pc_ffwd_init_impl(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* x, float* const* A_out,
                      const float* y, float* loss_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (desc && desc->free_input) return fail(JPCB_EINVAL, "init_activities_with_ffwd needs an input (free_input is set)");
  Run r;
  RET(r.init(desc, W, b, x, y, workspace, workspace_bytes, stream, 1));
  if (!A_out) return fail(JPCB_EINVAL, "null activity list");
  RET(r.zero_red());
  RET(r.prep_weights());
  RET(r.prep_inputs(nullptr));
  const bool want_loss = loss_out && y;
  RET(r.ffwd(A_out, want_loss ? r.ws.red + RED_LOSS : nullptr, false, true));
  if (want_loss) RET(r.finalize(nullptr, nullptr, false, loss_out));
  return JPCB_OK;
}

static int pc_energy_impl(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* const* A, const float* x,
                   const float* y, float* E_layers, void* workspace, size_t workspace_bytes, void* stream) {
  RET(split_free(desc, x, A));
  Run r;
  RET(r.init(desc, W, b, x, y, workspace, workspace_bytes, stream));
  if (!A || !y || !E_layers) return fail(JPCB_EINVAL, "null activities / target / output");
  RET(r.zero_red());
  RET(r.prep_weights());
  RET(r.prep_inputs(A));
  RET(r.errors(0, A, false, true));
  return r.finalize(E_layers, nullptr, false, nullptr);
}

static int pc_activity_grad_impl(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* const* A,
                          const float* x, const float* y, float* F, float* const* gA, void* workspace, size_t workspace_bytes,
                          void* stream) {
  if (!gA) return fail(JPCB_EINVAL, "null activities / target / output");
  float* gx = nullptr;
  if (desc && desc->free_input) { gx = gA[0]; gA = gA + 1; }
  RET(split_free(desc, x, A));
  Run r;
  RET(r.init(desc, W, b, x, y, workspace, workspace_bytes, stream));
  if (!A || !y || !gA) return fail(JPCB_EINVAL, "null activities / target / output");
  RET(r.zero_red());
  RET(r.prep_weights());
  RET(r.prep_inputs(A));
  RET(r.errors(0, A, false, true));
  r.x_out = gx;
  RET(r.grads(GRAD_OUT, r.invB, A, nullptr, gA, false, false));
  const size_t n_dummy = (size_t)r.Bl * desc->dims[r.L];
  scale_kernel<<<ew_grid(n_dummy, 256), 256, 0, r.st>>>(A[r.L - 1], gA[r.L - 1], n_dummy, desc->activity_decay * r.invB); count_launch();
  CK(cudaGetLastError());
  if (F) {
    const bool ad = desc->activity_decay != 0.f;
    if (ad) RET(r.sumsq_all(A, r.ws.red + RED_SUMSQ));
    RET(r.finalize(nullptr, F, ad, nullptr));
  }
  return JPCB_OK;
}

static int pc_infer_impl(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, float* const* A, const float* x,
                  const float* y, float* sumsq_out, void* workspace, size_t workspace_bytes, void* stream) {
  RET(split_free(desc, x, A));
  Run r;
  RET(r.init(desc, W, b, x, y, workspace, workspace_bytes, stream));
  if (!A || !y) return fail(JPCB_EINVAL, "null activities / target");
  RET(r.prep_weights());
  RET(r.prep_inputs(A));
  if (r.hoist) RET(r.predict_p1());
  RET(r.infer_loop(A));
  if (sumsq_out) {
    CK(cudaMemsetAsync(sumsq_out, 0, 2 * sizeof(float), r.st));
    RET(r.sumsq_all(A, sumsq_out));
    size_t numel = 0;
    for (int l = r.fr ? -1 : 0; l < r.L; ++l) numel += (size_t)r.Bl * desc->dims[l + 1];
    const float nf = (float)numel;
    CK(cudaMemcpyAsync(sumsq_out + 1, &nf, sizeof(float), cudaMemcpyHostToDevice, r.st));
  }
  return JPCB_OK;
}

int jpcb_pc_heun_trial(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* const* A, const float* x,
                       const float* y, float dt, float rtol, float atol, float* const* A_new, float* out2, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (!A_new) return fail(JPCB_EINVAL, "null activities / target / outputs");
  float* x_new = nullptr;
  if (desc && desc->free_input) { x_new = A_new[0]; A_new = A_new + 1; }
  RET(split_free(desc, x, A));
  Run r;
  RET(r.init(desc, W, b, x, y, workspace, workspace_bytes, stream));
  if (!A || !y || !A_new || !out2) return fail(JPCB_EINVAL, "null activities / target / outputs");
  if (desc->solver != JPCB_SOLVER_HEUN) return fail(JPCB_EINVAL, "jpcb_pc_heun_trial needs desc->solver == JPCB_SOLVER_HEUN");
  const int L = r.L;
  CK(cudaMemsetAsync(out2, 0, 2 * sizeof(float), r.st));
  RET(r.prep_weights());
  RET(r.prep_inputs(A));
  const bool fr = r.fr, hoist = r.hoist;
  if (hoist) RET(r.predict_p1());
  const float c = dt * r.invB;
  std::vector<float*> zt(r.ws.Zt.begin(), r.ws.Zt.begin() + (L - 1));
  zt.push_back(A_new[L - 1]);  // dummy slot is never read by the error tasks
  RET(r.errors(hoist ? 1 : 0, A, false, false));
  r.x_out = fr ? r.ws.Zx : nullptr;
  RET(r.grads(GRAD_HEUN1, c, A, nullptr, zt.data(), true, hoist));
  RET(r.errors(hoist ? 1 : 0, zt.data(), true, false));
  r.x_out = x_new;
  RET(r.grads(GRAD_HEUN2, c, zt.data(), A, A_new, false, hoist, 1, out2, rtol, atol));
  const size_t n_dummy = (size_t)r.Bl * desc->dims[L];
  if (desc->activity_decay != 0.f) {  // the dummy slot decays like every other entry of the list
    heun_dummy_kernel<<<ew_grid(n_dummy, 256), 256, 0, r.st>>>(A[L - 1], A_new[L - 1], n_dummy, c * desc->activity_decay, rtol, atol, out2);
    count_launch();
    CK(cudaGetLastError());
  } else {
    CK(cudaMemcpyAsync(A_new[L - 1], A[L - 1], n_dummy * sizeof(float), cudaMemcpyDeviceToDevice, r.st));
  }
  RET(r.sumsq_all(A_new, out2 + 1, x_new));
  return JPCB_OK;
}

static int pc_param_grads_impl(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* const* A,
                        const float* x, const float* y, float* const* gW, float* const* gb, float* E_layers, void* workspace,
                        size_t workspace_bytes, void* stream) {
  RET(split_free(desc, x, A));
  Run r;
  RET(r.init(desc, W, b, x, y, workspace, workspace_bytes, stream));
  if (!A || !y || !gW) return fail(JPCB_EINVAL, "null activities / target / output");
  RET(r.zero_red());
  RET(r.prep_weights());
  RET(r.prep_inputs(A));
  RET(r.errors(0, A, false, E_layers != nullptr));
  RET(r.wgrads(A, gW, gb));
  if (E_layers) RET(r.finalize(E_layers, nullptr, false, nullptr));
  return JPCB_OK;
}

int jpcb_hpc_param_grads(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* const* U,
                         const float* const* T, float* const* gW, float* const* gb, float* E_layers, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (!U || !T || !U[0]) return fail(JPCB_EINVAL, "null layer input / target lists");
  if (desc) {
    if (desc->free_input || desc->loss != JPCB_LOSS_MSE) return fail(JPCB_EINVAL, "layer-wise regression: mse, with an input");
    if (desc->n_layers >= 1 && desc->n_layers <= JPCB_MAX_LAYERS && any_post(desc)) return fail(JPCB_ENOTIMPL, "layer-wise regression: layers of the form act(W u + b) are not on this path");
    for (int l = 0; l < desc->n_layers && l < JPCB_MAX_LAYERS; ++l)
      if (desc->skip[l] || desc->scalings[l] != 1.f) return fail(JPCB_EINVAL, "layer-wise regression: no skips, unit scalings (layer %d)", l);
  }
  Run r;
  RET(r.init(desc, W, b, U[0], T[desc ? desc->n_layers - 1 : 0], workspace, workspace_bytes, stream));
  for (int l = 0; l < r.L; ++l)
    if (!U[l] || !T[l]) return fail(JPCB_EINVAL, "null input / target of layer %d", l);
  RET(r.zero_red());
  RET(r.prep_weights());
  RET(r.prep_inputs(U + 1));  // layer l >= 1 reads act_l(U[l])
  // e_l = 2 (T_l - f_l(U_l)): the energy sum e_l^2 / 4 = |T_l - f_l|^2 and the gradients -(1/B) e_l^T act_l(U_l) then
  // carry hpc_energy_fn's missing 1/2 (jpc/_core/_energies.py:228-243)
  RET(r.errors(0, U + 1, false, E_layers != nullptr, T, 2.f));
  if (gW) RET(r.wgrads(U + 1, gW, gb));
  if (E_layers) RET(r.finalize(E_layers, nullptr, false, nullptr));
  return JPCB_OK;
}

// ---------------------------------------------------------------------------------------------
// error-reparameterised PC: the state is the errors, the activities follow from them
// ---------------------------------------------------------------------------------------------
namespace {
struct EpcWs {
  std::vector<float*> Z;    // [L-1] z_1 .. z_{L-1} of the error-perturbed forward pass
  std::vector<float*> Rho;  // [L]   tensor-core path: fp32 back-propagated signal rho_{l+1} (the fp32 path keeps it in ws.err)
  float* R0 = nullptr;      // free_input: rho_0 [Bl, d_0], the signal arriving at z_0
  size_t bytes = 0;
};
void carve_epc(const jpcb_pc_desc* d, void* base, size_t pc_bytes, EpcWs* w) {
  const int L = d->n_layers;
  const size_t Bl = (size_t)d->batch_local;
  Bump b{reinterpret_cast<char*>(base)};
  b.off = pc_bytes;
  w->Z.assign(L, nullptr);
  w->Rho.assign(L, nullptr);
  for (int l = 0; l + 1 < L; ++l) w->Z[l] = b.take<float>(Bl * d->dims[l + 1]);
  if (d->dtype == JPCB_DTYPE_BF16)  // (the split path keeps fp32 errors in ws.err like the fp32 path)
    for (int l = 0; l < L; ++l) w->Rho[l] = b.take<float>(Bl * d->dims[l + 1]);
  if (d->free_input) w->R0 = b.take<float>(Bl * d->dims[0]);
  w->bytes = ((b.off + 1023) & ~(size_t)1023) + 1024;
}
}  // namespace

size_t jpcb_epc_workspace_bytes(const jpcb_pc_desc* desc) {
  if (validate(desc) != JPCB_OK) return 0;
  Ws w;
  carve(desc, nullptr, &w);
  EpcWs x;
  carve_epc(desc, nullptr, w.bytes, &x);
  return x.bytes;
}

int jpcb_epc_grads(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* const* E, const float* x,
                   const float* y, float* F, float* const* gE, float* const* gW, float* const* gb, void* workspace,
                   size_t workspace_bytes, void* stream) {
  // free_input (x = None in the reference, jpc/_core/_energies.py:429-432): errors[0] plays z_0; the lists have L + 1 entries
  const bool fr = desc && desc->free_input;
  float* gx = nullptr;
  if (fr && gE) { gx = gE[0]; gE = gE + 1; }
  RET(split_free(desc, x, E));
  Run r;
  RET(r.init(desc, W, b, x, y, workspace, workspace_bytes, stream));
  if (!E || !y) return fail(JPCB_EINVAL, "null errors / target");
  if (desc->activity_decay != 0.f || desc->output_energy_scaling != 1.f)
    return fail(JPCB_EINVAL, "epc_energy_fn has no activity decay / output energy scaling");
  if (r.post) return fail(JPCB_ENOTIMPL, "ePC: layers of the form act(W u + b) are not on this path");
  EpcWs X;
  carve_epc(desc, workspace, r.ws.bytes, &X);
  if (workspace_bytes < X.bytes) return fail(JPCB_EWORKSPACE, "workspace too small: need %zu bytes (jpcb_epc_workspace_bytes), got %zu", X.bytes, workspace_bytes);
  const int L = r.L, Bl = r.Bl;
  const bool tc = r.tc, ce = desc->loss == JPCB_LOSS_CE;
  auto rho = [&](int l) { return (tc && !r.x3) ? X.Rho[l] : r.ws.err[l]; };  // rho_{l+1}, fp32
  RET(r.zero_red());
  RET(r.prep_weights());
  RET(r.prep_inputs(nullptr));
  // ---- error-perturbed forward pass (jpc/_core/_energies.py:434-447): z_{l+1} = a_l f_l(z_l) + eps_{l+1} (+ z_l) ----
  for (int l = 0; l + 1 < L; ++l) {
    GTask t;
    r.fwd_operands(t, l, nullptr, false);
    t.e.mode = EPI_FFWD;
    t.e.skip = 1.f; t.e.U = E[l];
    if (desc->skip[l]) t.e.Gin = l == 0 ? x : X.Z[l - 1];  // any non-None skip below the output layer (:441-442)
    t.e.O = X.Z[l];
    t.e.act = desc->act[l + 1];
    if (tc) r.set_ob(t, r.ws.Hb[l + 1]); else t.e.Oh = r.ws.Hf[l + 1];
    RET(r.launch({t}));
  }
  {  // output layer: z_L = a f(z_{L-1}) - eps_L; residual rho_L = y - z_L (or the softmax error) and its energy
    GTask t;
    r.fwd_operands(t, L - 1, nullptr, false);
    t.e.skip = -1.f; t.e.U = E[L - 1];
    if (ce) {
      t.e.mode = EPI_FFWD;
      t.e.O = r.ws.logits;
      RET(r.launch({t}));
      RET(r.ce_rows(r.ws.logits, 1.f, rho(L - 1), r.ws.Eb[L - 1], r.ws.red + (L - 1)));
    } else {
      t.e.mode = EPI_ERR;
      t.e.T = y; t.e.escale = 1.f;
      t.e.O = rho(L - 1); r.set_ob(t, r.ws.Eb[L - 1]);
      t.e.red = r.ws.red + (L - 1);
      RET(r.launch({t}));
    }
  }
  // ---- energy: (sum of 1/2 |.|^2 over the FIRST L-1 entries of the caller's list + output term) / B  (:449-457).
  // Supervised: eps_1 .. eps_{L-1}.  x = None: the list starts with z_0, so z_0 and eps_1 .. eps_{L-2} (the reference's
  // loop bound does not move: eps_{L-1} goes unpenalised -- kept as is) ----
  const int n_pen = fr ? L - 2 : L - 1;  // penalised entries among E[0..] = eps_1 ..
  if (F) {
    if (fr) {
      const size_t n = (size_t)Bl * desc->dims[0];
      sumsq_kernel<<<ew_grid(n, 1024), 256, 0, r.st>>>(x, n, r.ws.red + RED_SUMSQ); count_launch();
    }
    for (int l = 0; l < n_pen; ++l) {
      const size_t n = (size_t)Bl * desc->dims[l + 1];
      sumsq_kernel<<<ew_grid(n, 1024), 256, 0, r.st>>>(E[l], n, r.ws.red + RED_SUMSQ); count_launch();
    }
    CK(cudaGetLastError());
    finalize_kernel<<<1, 32, 0, r.st>>>(r.ws.red, L, r.invB, nullptr, F, r.ws.red + RED_SUMSQ, 1.f, r.ws.red + RED_LOSS, nullptr);
    count_launch();
    CK(cudaGetLastError());
  }
  if (!gE && !gW) return JPCB_OK;
  // ---- backward pass: rho_l = act_l'(z_l) . (a_l rho_{l+1} W_l) + s_l rho_{l+1};  dF/d eps_l = (eps_l - rho_l) / B ----
  // (the PC activity-gradient epilogue with "own error" 0 and coefficient -1 returns exactly rho_l)
  for (int l = L - 1; l >= 1; --l) {
    GTask t;
    t.e.mode = EPI_GRAD; t.e.sub = GRAD_OUT; t.e.act = desc->act[l];
    t.e.M = Bl; t.e.N = desc->dims[l]; t.K = desc->dims[l + 1];
    t.a_p = rho(l); t.a_smn = desc->dims[l + 1]; t.a_sk = 1;
    t.b_p = r.ws.WTf[l]; t.b_smn = desc->dims[l + 1]; t.b_sk = 1;
    t.a_b = r.ws.Eb[l]; t.b_b = r.ws.WTb[l];
    r.split_geom(t, desc->dims[l + 1], desc->dims[l + 1]);
    t.e.alpha = desc->scalings[l];
    t.e.skip = (desc->skip[l] && l + 1 < L) ? 1.f : 0.f;  // (the output layer's forward pass has no skip term)
    t.e.En = rho(l); t.e.Enb = r.x3 ? nullptr : r.ws.Eb[l];
    t.e.Z = X.Z[l - 1]; t.e.P = X.Z[l - 1];
    t.e.c = -1.f; t.e.gscale = 1.f;
    t.e.O = rho(l - 1);
    RET(r.launch({t}));
    if (tc) {
      r.operand_copy(rho(l - 1), r.ws.Eb[l - 1], desc->dims[l], ACT_LINEAR);
      CK(cudaGetLastError());
    }
  }
  if (fr && gE) {  // the signal arriving at z_0 = errors[0]: rho_0 = act_0'(z_0) . (a_0 rho_1 W_0) + s_0 rho_1
    GTask t;
    t.e.mode = EPI_GRAD; t.e.sub = GRAD_OUT; t.e.act = desc->act[0];
    t.e.M = Bl; t.e.N = desc->dims[0]; t.K = desc->dims[1];
    t.a_p = rho(0); t.a_smn = desc->dims[1]; t.a_sk = 1;
    t.b_p = r.ws.WTf[0]; t.b_smn = desc->dims[1]; t.b_sk = 1;
    t.a_b = r.ws.Eb[0]; t.b_b = r.ws.WTb[0];
    r.split_geom(t, desc->dims[1], desc->dims[1]);
    t.e.alpha = desc->scalings[0];
    t.e.skip = (desc->skip[0] && L > 1) ? 1.f : 0.f;
    t.e.En = rho(0); t.e.Enb = r.x3 ? nullptr : r.ws.Eb[0];
    t.e.Z = x; t.e.P = x;
    t.e.c = -1.f; t.e.gscale = 1.f;
    t.e.O = X.R0;
    RET(r.launch({t}));
  }
  if (gE) {
    if (fr) {
      if (!gx) return fail(JPCB_EINVAL, "null error-gradient buffer 0");
      const size_t n = (size_t)Bl * desc->dims[0];
      axpby_kernel<<<ew_grid(n, 256), 256, 0, r.st>>>(x, X.R0, gx, n, r.invB, -r.invB); count_launch();
    }
    for (int l = 0; l < L; ++l) {
      if (!gE[l]) return fail(JPCB_EINVAL, "null error-gradient buffer %d", l);
      const size_t n = (size_t)Bl * desc->dims[l + 1];
      // penalised errors: (eps - rho) / B; unpenalised hidden ones (x = None: eps_{L-1}): -rho / B; the output "error"
      // enters z_L with a minus sign and carries no quadratic term: rho_L / B
      if (l < n_pen) axpby_kernel<<<ew_grid(n, 256), 256, 0, r.st>>>(E[l], rho(l), gE[l], n, r.invB, -r.invB);
      else scale_kernel<<<ew_grid(n, 256), 256, 0, r.st>>>(rho(l), gE[l], n, l + 1 < L ? -r.invB : r.invB);
      count_launch();
    }
    CK(cudaGetLastError());
  }
  if (gW) RET(r.wgrads(nullptr, gW, gb));  // -(a_l / B) rho_{l+1}^T act_l(z_l): the PC weight-gradient schedule on rho
  return JPCB_OK;
}

static int pc_train_step_impl(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* x, const float* y,
                       float* const* A_out, float* loss_out, float* E_layers, float* const* gW, float* const* gb, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (desc && desc->free_input)
    return fail(JPCB_ENOTIMPL, "the fused train step starts from the feed-forward init, which needs an input; "
                               "unsupervised steps are jpcb_pc_infer + jpcb_pc_param_grads on caller-initialised activities");
  Run r;
  RET(r.init(desc, W, b, x, y, workspace, workspace_bytes, stream));
  if (!A_out || !y) return fail(JPCB_EINVAL, "null activities / target");
  RET(r.zero_red());
  RET(r.prep_weights());
  RET(r.prep_inputs(nullptr));
  // With the wavefront schedule the layers the front has not reached after T steps keep EXACTLY zero errors (their buffers
  // were zeroed and are never written): on the fp32 path the error pass at the solution skips them too -- zero energies and
  // zero gradients by construction, whatever order the feed-forward kernel summed in (so the chain kernel may be used).
  const bool wave = r.wavefront_ok(), wave_tail = wave && !r.tc;
  // (the hoisted first-layer prediction P1 is the ENERGY's: no skip term at layer 0, jpc/_core/_energies.py:111-126; the init
  //  honours a skip there, _init.py:69-78 -- with one, z_1 is not P1 and P1 takes its own GEMM)
  const bool p1_from_init = r.hoist && !desc->skip[0];
  RET(r.ffwd(A_out, r.ws.red + RED_LOSS, p1_from_init, wave_tail));   // init + loss at the init + hoisted P1 + operand copies
  if (r.hoist && !p1_from_init) RET(r.predict_p1());
  RET(r.infer_loop(A_out, true));
  const int reach = r.L - 1 - desc->n_steps;  // lowest layer whose error can be non-zero at the solution
  RET(r.errors(r.hoist ? (wave_tail && reach > 1 ? reach : 1) : 0, A_out, false, true));  // errors + layer energies at the solution
  if (r.hoist && !(wave_tail && reach > 0)) RET(r.error0_from_p1(A_out, true));
  if (gW) RET(r.wgrads(A_out, gW, gb));           // (NULL: the caller follows up with jpcb_pc_param_grads_range)
  return r.finalize(E_layers, nullptr, false, loss_out);
}

int jpcb_pc_param_grads_range(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* x, int layer_lo,
                              int layer_hi, float* const* gW, float* const* gb, void* workspace, size_t workspace_bytes, void* stream) {
  if (desc && desc->free_input) return fail(JPCB_ENOTIMPL, "jpcb_pc_param_grads_range follows jpcb_pc_train_step, which needs an input");
  Run r;
  RET(r.init(desc, W, b, x, nullptr, workspace, workspace_bytes, stream));
  if (!gW) return fail(JPCB_EINVAL, "null gradient list");
  if (layer_lo < 0 || layer_hi > r.L || layer_lo >= layer_hi) return fail(JPCB_EINVAL, "bad layer range [%d, %d) of %d layers", layer_lo, layer_hi, r.L);
  return r.wgrads(nullptr, gW, gb, layer_lo, layer_hi);
}


// ---- planned entry points: identical calls are captured into a CUDA graph once and replayed (plan.cu) ----
namespace {
// list length to key on; < 0: the descriptor is unusable and the body reports why
inline int key_layers(const jpcb_pc_desc* d, int min_layers) {
  return (d && d->n_layers >= min_layers && d->n_layers <= JPCB_MAX_LAYERS) ? d->n_layers : -1;
}
inline void key_net(PlanKey& k, const jpcb_pc_desc* d, const float* const* W, const float* const* b, const void* ws, size_t ws_bytes) {
  k.bytes(d, sizeof(*d));
  k.list(W, d->n_layers);
  k.list(d->has_bias ? b : nullptr, d->n_layers);
  k.ptr(ws);
  k.word(ws_bytes);
}
}  // namespace

int jpcb_pc_ffwd_init(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* x, float* const* A_out,
                      const float* y, float* loss_out, void* workspace, size_t workspace_bytes, void* stream) {
  auto body = [&](void* s) { return pc_ffwd_init_impl(desc, W, b, x, A_out, y, loss_out, workspace, workspace_bytes, s); };
  const int L = key_layers(desc, 1);
  if (L < 0 || !W || !A_out) return body(stream);
  PlanKey k(0x7001);
  key_net(k, desc, W, b, workspace, workspace_bytes);
  k.ptr(x); k.list(A_out, L); k.ptr(y); k.ptr(loss_out);
  return plan_run(k, stream, body);
}

int jpcb_pc_energy(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* const* A, const float* x,
                   const float* y, float* E_layers, void* workspace, size_t workspace_bytes, void* stream) {
  auto body = [&](void* s) { return pc_energy_impl(desc, W, b, A, x, y, E_layers, workspace, workspace_bytes, s); };
  const int L = key_layers(desc, 2);
  if (L < 0 || !W || !A) return body(stream);
  PlanKey k(0x7002);
  key_net(k, desc, W, b, workspace, workspace_bytes);
  k.list(A, L + (desc->free_input ? 1 : 0)); k.ptr(x); k.ptr(y); k.ptr(E_layers);
  return plan_run(k, stream, body);
}

int jpcb_pc_activity_grad(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* const* A,
                          const float* x, const float* y, float* F, float* const* gA, void* workspace, size_t workspace_bytes,
                          void* stream) {
  auto body = [&](void* s) { return pc_activity_grad_impl(desc, W, b, A, x, y, F, gA, workspace, workspace_bytes, s); };
  const int L = key_layers(desc, 2);
  if (L < 0 || !W || !A || !gA) return body(stream);
  PlanKey k(0x7003);
  key_net(k, desc, W, b, workspace, workspace_bytes);
  const int n = L + (desc->free_input ? 1 : 0);
  k.list(A, n); k.ptr(x); k.ptr(y); k.ptr(F); k.list(gA, n);
  return plan_run(k, stream, body);
}

int jpcb_pc_infer(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, float* const* A, const float* x,
                  const float* y, float* sumsq_out, void* workspace, size_t workspace_bytes, void* stream) {
  auto body = [&](void* s) { return pc_infer_impl(desc, W, b, A, x, y, sumsq_out, workspace, workspace_bytes, s); };
  const int L = key_layers(desc, 2);
  // (sumsq_out: the call copies a host scalar to the device -- not capturable)
  if (L < 0 || !W || !A || sumsq_out) return body(stream);
  PlanKey k(0x7004);
  key_net(k, desc, W, b, workspace, workspace_bytes);
  k.list(A, L + (desc->free_input ? 1 : 0)); k.ptr(x); k.ptr(y);
  return plan_run(k, stream, body);
}

int jpcb_pc_param_grads(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* const* A,
                        const float* x, const float* y, float* const* gW, float* const* gb, float* E_layers, void* workspace,
                        size_t workspace_bytes, void* stream) {
  auto body = [&](void* s) { return pc_param_grads_impl(desc, W, b, A, x, y, gW, gb, E_layers, workspace, workspace_bytes, s); };
  const int L = key_layers(desc, 2);
  if (L < 0 || !W || !A || !gW) return body(stream);
  PlanKey k(0x7005);
  key_net(k, desc, W, b, workspace, workspace_bytes);
  k.list(A, L + (desc->free_input ? 1 : 0)); k.ptr(x); k.ptr(y); k.list(gW, L); k.list(desc->has_bias ? gb : nullptr, L); k.ptr(E_layers);
  return plan_run(k, stream, body);
}

int jpcb_pc_train_step(const jpcb_pc_desc* desc, const float* const* W, const float* const* b, const float* x, const float* y,
                       float* const* A_out, float* loss_out, float* E_layers, float* const* gW, float* const* gb, void* workspace,
                       size_t workspace_bytes, void* stream) {
  auto body = [&](void* s) { return pc_train_step_impl(desc, W, b, x, y, A_out, loss_out, E_layers, gW, gb, workspace, workspace_bytes, s); };
  const int L = key_layers(desc, 2);
  if (L < 0 || !W || !A_out) return body(stream);
  PlanKey k(0x7006);
  key_net(k, desc, W, b, workspace, workspace_bytes);
  k.ptr(x); k.ptr(y); k.list(A_out, L); k.ptr(loss_out); k.ptr(E_layers); k.list(gW, L); k.list((desc->has_bias && gW) ? gb : nullptr, L);
  return plan_run(k, stream, body);
}

}  // extern "C"
