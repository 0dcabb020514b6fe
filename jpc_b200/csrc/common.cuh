// common.cuh -- device-side vocabulary shared by the fp32 CUDA-core kernels (simt_gemm.cu) and the
// tcgen05 tensor-core kernels (tc_gemm.cu): activation functions with the reference's (jax.nn)
// conventions, the fused-epilogue descriptor, and the epilogue itself.
//
// Every GEMM on the PC hot path is  D[M,N] = A[M,K] * B[N,K]^T  followed by ONE of four fused
// epilogues (reference math in jpc/_core/_energies.py:103-126 and its
// reverse-mode derivative jpc/_core/_grads.py:79-91,487-499):
//   FFWD  z_out = a (D + b) + s u                        (feed-forward init, _init.py:68-80)
//   ERR   e     = lam (T - a (D + b) - s u),  E += lam/2 (..)^2   (prediction errors + layer energy)
//   GRAD  g     = Gin + w (e_l - act'(z) (a D) - s e_next) + ad z ; then Euler/Heun update of z
//   SCALE out   = a D                                    (weight gradients, K = batch)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace jpcb {

enum : int { ACT_LINEAR = 0, ACT_TANH, ACT_HARD_TANH, ACT_RELU, ACT_LEAKY_RELU, ACT_GELU, ACT_SELU, ACT_SILU };
enum : int { EPI_FFWD = 0, EPI_ERR = 1, EPI_GRAD = 2, EPI_SCALE = 3 };
enum : int { GRAD_OUT = 0, GRAD_EULER = 1, GRAD_HEUN1 = 2, GRAD_HEUN2 = 3 };

#define JPCB_SELU_ALPHA 1.6732632423543772848170429916717f
#define JPCB_SELU_SCALE 1.0507009873554804934193349852946f

// FAST=false: accurate libm-grade fp32 (1e-5 parity path).  FAST=true: MUFU approximations
// (tanh.approx / ex2.approx) for the bf16 tensor-core path, whose operands are rounded to 8 bits anyway.
template <bool FAST>
__device__ __forceinline__ float tanh_f(float x) {
  if constexpr (FAST) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  } else {
    return tanhf(x);
  }
}
template <bool FAST>
__device__ __forceinline__ float sigmoid_f(float x) {
  if constexpr (FAST) return __fdividef(1.f, 1.f + __expf(-x));
  else return 1.f / (1.f + expf(-x));
}

// act(x) and act'(x); conventions at kinks follow jax.nn.
template <bool FAST>
__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case ACT_LINEAR: return x;
    case ACT_TANH: return tanh_f<FAST>(x);
    case ACT_HARD_TANH: return x > 1.f ? 1.f : (x < -1.f ? -1.f : x);
    case ACT_RELU: return x > 0.f ? x : 0.f;
    case ACT_LEAKY_RELU: return x >= 0.f ? x : 0.01f * x;
    case ACT_GELU: {
      const float c = 0.7978845608028654f;
      float u = c * (x + 0.044715f * x * x * x);
      return 0.5f * x * (1.f + tanh_f<FAST>(u));
    }
    case ACT_SELU: return JPCB_SELU_SCALE * (x > 0.f ? x : JPCB_SELU_ALPHA * (FAST ? (__expf(x) - 1.f) : expm1f(x)));
    case ACT_SILU: return x * sigmoid_f<FAST>(x);
  }
  return x;
}
template <bool FAST>
__device__ __forceinline__ float act_bwd(int act, float x) {
  switch (act) {
    case ACT_LINEAR: return 1.f;
    case ACT_TANH: { float t = tanh_f<FAST>(x); return 1.f - t * t; }
    case ACT_HARD_TANH: return (x > 1.f || x < -1.f) ? 0.f : 1.f;
    case ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case ACT_LEAKY_RELU: return x >= 0.f ? 1.f : 0.01f;
    case ACT_GELU: {
      const float c = 0.7978845608028654f;
      float x2 = x * x;
      float t = tanh_f<FAST>(c * (x + 0.044715f * x * x2));
      return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * c * (1.f + 3.f * 0.044715f * x2);
    }
    case ACT_SELU: return JPCB_SELU_SCALE * (x > 0.f ? 1.f : JPCB_SELU_ALPHA * (FAST ? __expf(x) : expf(x)));
    case ACT_SILU: { float s = sigmoid_f<FAST>(x); return s * (1.f + x * (1.f - s)); }
  }
  return 1.f;
}

// Fused-epilogue descriptor: plain data, passed to kernels by value inside the grouped-launch
// parameter block.  All [M,N] tensors are dense row-major with row pitch N.
struct Epi {
  int mode;   // EPI_*
  int sub;    // GRAD_* when mode == EPI_GRAD
  int act;    // activation id: act'(z) in GRAD, act(out) for the bf16 operand copy `Ob`
  int M, N;
  int accum;  // SCALE: out += a*D instead of =
  int pre_act;  // FFWD / ERR: activation applied to (D + b) before the scaling -- layers of the form act(W u + b) (Form B);
                // ERR then writes the RAW error to O and act'(D + b) * error to the operand copies Ob / Oh.  0 = none.
                // Honoured by the generic (descriptor-driven) instantiation only: the host never routes such a task to a
                // specialised variant.
  // `Ob` as a 3-way split copy (fp32-parity tensor-core path): v = hi + mid + lo in bf16, stored at
  // Ob[row * ob_ld + p * ob_part + col], p = 0, 1, 2.  ob_part == 0: plain bf16 copy with row pitch N.
  int ob_ld, ob_part;
  float alpha;   // layer scaling a_l (SCALE: -(a_l/B))
  float skip;    // s_l (0 or 1)
  float escale;  // ERR: lam (output_energy_scaling on the output layer, else 1)
  float c;       // GRAD: OUT -> multiplier of g (1/B); EULER/HEUN -> dt/B
  float ad;      // GRAD: activity_decay
  float gscale;  // GRAD: weight w of this term (bPC alpha_f / alpha_b; 1 for PC)
  float tol_r, tol_a;  // GRAD_HEUN2 with `red`: rtol / atol of the embedded-error norm (adaptive stepping)
  const float* bias;           // [N] or null
  const float* T;              // ERR: target
  const float* U;              // FFWD/ERR: skip input
  const float* Z;              // GRAD: state at which act' is evaluated (z, or Heun stage state)
  const float* Z0;             // GRAD_HEUN2: state at the start of the step
  const float* El;             // GRAD: e_l fp32, or
  const __nv_bfloat16* Elb;    //       e_l bf16, or (both null)
  const float* P;              //       e_l = Z - P  (hoisted layer-0 prediction, or clamped-side target)
  const float* En;             // GRAD: e_{l+1} for the skip term (fp32) or
  const __nv_bfloat16* Enb;    //       bf16
  const float* Gin;            // GRAD: partial gradient to add (bPC second direction) or null
  float* G1;                   // GRAD_HEUN1: out g1;  GRAD_HEUN2: in g1
  float* O;                    // fp32 out (z_out / e / g / scaled D)
  float* O2;                   // FFWD: optional second fp32 copy
  __nv_bfloat16* Ob;           // bf16 out: act(z_out) (FFWD / GRAD updates) or e (ERR)
  float* Oh;                   // fp32 out: act(z_out) (FFWD / GRAD updates) -- the fp32 path's GEMM operand copy
  float* red;                  // ERR: energy accumulator; FFWD: loss accumulator (needs Y)
  const float* Y;              // FFWD: loss target
};

// Copies a descriptor out of kernel-parameter space into registers and makes every field opaque to
// the compiler, so later uses read registers.  Descriptors sit in a runtime-indexed array of the
// grouped-launch parameter block; a runtime-indexed constant load is a variable-latency operation
// that shares scoreboard slots with outstanding global loads, so re-reading a field in the middle
// of the epilogue's load stream would wait for every load in flight.
__device__ __forceinline__ Epi epi_pin(const Epi& src) {
  Epi e = src;
  asm volatile("" : "+r"(e.mode), "+r"(e.sub), "+r"(e.act), "+r"(e.M), "+r"(e.N), "+r"(e.accum), "+r"(e.pre_act), "+r"(e.ob_ld), "+r"(e.ob_part));
  asm volatile("" : "+f"(e.alpha), "+f"(e.skip), "+f"(e.escale), "+f"(e.c), "+f"(e.ad), "+f"(e.gscale), "+f"(e.tol_r), "+f"(e.tol_a));
  asm volatile("" : "+l"(e.bias), "+l"(e.T), "+l"(e.U), "+l"(e.Z), "+l"(e.Z0), "+l"(e.El), "+l"(e.Elb), "+l"(e.P));
  asm volatile("" : "+l"(e.En), "+l"(e.Enb), "+l"(e.Gin), "+l"(e.G1), "+l"(e.O), "+l"(e.O2), "+l"(e.Ob), "+l"(e.Oh), "+l"(e.red), "+l"(e.Y));
  return e;
}

__device__ __forceinline__ void ld4(const float* __restrict__ p, uint32_t idx, int nvalid, bool vec, float (&v)[4]) {
  if (vec) {
    // __ldcg / __stcg: explicit GLOBAL-space accesses (the descriptor pointers lose their address space in
    // epi_pin, and generic LD/ST cost extra descriptor moves), cached at L2 only -- these are streaming tensors
    float4 t = __ldcg(reinterpret_cast<const float4*>(p + idx));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = j < nvalid ? p[idx + j] : 0.f;
  }
}
__device__ __forceinline__ void ld4(const __nv_bfloat16* __restrict__ p, uint32_t idx, int nvalid, bool vec, float (&v)[4]) {
  if (vec) {
    uint2 t = __ldcg(reinterpret_cast<const uint2*>(p + idx));
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = j < nvalid ? __bfloat162float(p[idx + j]) : 0.f;
  }
}
__device__ __forceinline__ void st4(float* __restrict__ p, uint32_t idx, int nvalid, bool vec, const float (&v)[4]) {
  if (vec) {
    __stcg(reinterpret_cast<float4*>(p + idx), make_float4(v[0], v[1], v[2], v[3]));
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (j < nvalid) p[idx + j] = v[j];
  }
}
__device__ __forceinline__ void st4(__nv_bfloat16* __restrict__ p, uint32_t idx, int nvalid, bool vec, const float (&v)[4]) {
  if (vec) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    __stcg(reinterpret_cast<uint2*>(p + idx), t);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (j < nvalid) p[idx + j] = __float2bfloat16_rn(v[j]);
  }
}

// bf16 operand copy of 4 values: plain (row pitch N), or the 3-way split v = hi + mid + lo of the fp32-parity path
// SPLIT = false: the instantiation never sees split copies (the hot bf16 variants stay free of this code)
template <bool SPLIT>
__device__ __forceinline__ void st_ob(const Epi& e, int row, int col, uint32_t idx, int nvalid, bool vec, const float (&v)[4]) {
  if constexpr (!SPLIT) {
    st4(e.Ob, idx, nvalid, vec, v);
  } else {
    if (e.ob_part == 0) {
      st4(e.Ob, idx, nvalid, vec, v);
      return;
    }
    float hi[4], mid[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hi[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
      const float r1 = v[j] - hi[j];
      mid[j] = __bfloat162float(__float2bfloat16_rn(r1));
      lo[j] = r1 - mid[j];
    }
    const uint32_t b0 = (uint32_t)row * (uint32_t)e.ob_ld + (uint32_t)col;
    st4(e.Ob, b0, nvalid, vec, hi);
    st4(e.Ob, b0 + (uint32_t)e.ob_part, nvalid, vec, mid);
    st4(e.Ob, b0 + 2u * (uint32_t)e.ob_part, nvalid, vec, lo);
  }
}

// The fused epilogue is split in two so that callers can put many independent global loads in
// flight before they consume any of them (the tensor-core epilogue issues a whole 32x32 patch of
// loads, then reads the accumulator from TMEM, then finishes):
//   epi_preload : issues the mode's per-element input loads for D[row, col..col+3] (raw values)
//   epi_finish  : consumes them + the 4 accumulator values, stores, returns the reduction term
struct EpiPre {
  float a[4];  // GRAD: z          ERR: target T    FFWD: loss target Y   SCALE(accum): old O
  float b[4];  // GRAD: e_l (fp32) or P; PLAIN == 2: the partial-gradient input Gin
  float c[4];  // GRAD, ELSRC == 2 (fp32 e_l next to a Gin in b)
  uint2 h;     // GRAD: e_l as 4 packed bf16 when Elb is the source.  Deliberately NOT overlaid on b: two
               // differently predicated loads into the same registers make the second one wait for
               // every load that shares the first one's scoreboard slot.
};

__device__ __forceinline__ void unpack_bf16x4(const uint2& t, float (&v)[4]) {
  __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}

// MODE / SUB / ACT >= 0 fix the corresponding descriptor field at compile time (the tensor-core
// epilogue dispatches once per tile so that its inner loop is straight-line code); -1 = read it
// from the descriptor at run time.
// ELSRC fixes where GRAD takes e_l from: 0 = Elb (packed bf16), 1 = z - P (hoisted prediction), 2 = El (fp32: the
// fp32-parity split path), -1 = run time.
// PLAIN == 2 (ELSRC == 0 or 2): the partial-gradient input Gin is part of the task and is loaded here into p.b.
template <int MODE = -1, int ELSRC = -1, int PLAIN = 0>
__device__ __forceinline__ void epi_preload(const Epi& e, int row, int col, int nvalid, bool vec, EpiPre& p) {
  const uint32_t idx = (uint32_t)row * (uint32_t)e.N + (uint32_t)col;  // M*N < 2^31 (validated on the host)
  const int mode = MODE >= 0 ? MODE : e.mode;
  if (mode == EPI_GRAD) {
    ld4(e.Z, idx, nvalid, vec, p.a);
    if (ELSRC == 0) {
      p.h = __ldcg(reinterpret_cast<const uint2*>(e.Elb + idx));
      if (PLAIN == 2) ld4(e.Gin, idx, nvalid, vec, p.b);
    } else if (ELSRC == 1) ld4(e.P, idx, nvalid, vec, p.b);
    else if (ELSRC == 2) {
      ld4(e.El, idx, nvalid, vec, p.c);
      if (PLAIN == 2) ld4(e.Gin, idx, nvalid, vec, p.b);
    } else if (e.El || !e.Elb) ld4(e.El ? e.El : e.P, idx, nvalid, vec, p.b);
    else if (vec) p.h = __ldcg(reinterpret_cast<const uint2*>(e.Elb + idx));
    else ld4(e.Elb, idx, nvalid, false, p.b);
  } else if (mode == EPI_ERR) {
    ld4(e.T, idx, nvalid, vec, p.a);  // (the skip input U, rare, is read in epi_finish)
  } else if (mode == EPI_FFWD) {
    if (e.red) ld4(e.Y, idx, nvalid, vec, p.a);
  } else if (e.accum) {
    ld4(e.O, idx, nvalid, vec, p.a);
  }
}

// `nvalid` (1..4) columns are inside the matrix; `vec` says 16-byte vector access is legal
// (N % 4 == 0 and nvalid == 4).  Returns this thread's contribution to the reduction (`red`).
// `have_bias4`: the caller already holds the 4 bias values of columns col..col+3 in `bias4`.
// PLAIN >= 1: the task has no skip term, no bias, unit output-energy weight and no activity decay -- the common
// case, checked on the host -- so those branches compile away; PLAIN == 1: no partial-gradient input either;
// PLAIN == 2: a partial-gradient input Gin, already loaded by epi_preload (bPC: the other direction's half).
template <bool FAST, int MODE = -1, int SUB = -1, int ACT = -1, int ELSRC = -1, int PLAIN = 0>
__device__ __forceinline__ float epi_finish(const Epi& e, int row, int col, const float (&acc)[4], int nvalid, bool vec,
                                            const EpiPre& p, bool have_bias4, const float (&bias4)[4]) {
  const uint32_t idx = (uint32_t)row * (uint32_t)e.N + (uint32_t)col;  // M*N < 2^31 (validated on the host)
  static_assert(PLAIN == 0 || MODE != EPI_GRAD || SUB == GRAD_EULER || SUB == GRAD_OUT, "plain GRAD variants exist for OUT / EULER only");
  // split (fp32-parity) operand copies reach the generic instantiation and the variants marked with ELSRC == 2 only
  constexpr bool SPLIT = MODE < 0 || ELSRC == 2;
  const int mode = MODE >= 0 ? MODE : e.mode;
  const int sub = SUB >= 0 ? SUB : e.sub;
  const int act = ACT >= 0 ? ACT : e.act;
  float r = 0.f;
  float out[4];
  if (mode == EPI_SCALE) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = (e.accum ? p.a[j] : 0.f) + e.alpha * acc[j];
    st4(e.O, idx, nvalid, vec, out);
    return 0.f;
  }
  if (mode == EPI_FFWD || mode == EPI_ERR) {
    float pred[4];
    const int pre_act = MODE < 0 ? e.pre_act : 0;  // (compiled into the generic instantiation only)
    float dpre[4] = {1.f, 1.f, 1.f, 1.f};          // post'(pre) of a Form B layer
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (PLAIN) {
        pred[j] = e.alpha * acc[j];
      } else {
        float bj = have_bias4 ? bias4[j] : ((e.bias != nullptr && j < nvalid) ? e.bias[col + j] : 0.f);
        float pre = acc[j] + bj;
        if (MODE < 0 && pre_act != ACT_LINEAR) {
          if (mode == EPI_ERR) dpre[j] = act_bwd<FAST>(pre_act, pre);
          pre = act_fwd<FAST>(pre_act, pre);
        }
        pred[j] = e.alpha * pre;
      }
    }
    if (!PLAIN && e.skip != 0.f) {
      float u[4];
      ld4(e.U, idx, nvalid, vec, u);
#pragma unroll
      for (int j = 0; j < 4; ++j) pred[j] += e.skip * u[j];
    }
    if (!PLAIN && mode == EPI_FFWD && e.Gin != nullptr) {  // a second additive input (ePC: error perturbation AND skip)
      float gi[4];
      ld4(e.Gin, idx, nvalid, vec, gi);
#pragma unroll
      for (int j = 0; j < 4; ++j) pred[j] += gi[j];
    }
    if (mode == EPI_FFWD) {
      st4(e.O, idx, nvalid, vec, pred);
      if (e.O2) st4(e.O2, idx, nvalid, vec, pred);
      if (e.Ob || e.Oh) {
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = act_fwd<FAST>(act, pred[j]);
        if (e.Ob) st_ob<SPLIT>(e, row, col, idx, nvalid, vec, out);
        if (e.Oh) st4(e.Oh, idx, nvalid, vec, out);
      }
      if (e.red) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < nvalid) { float d = p.a[j] - pred[j]; r += 0.5f * d * d; }
      }
    } else {  // EPI_ERR
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = j < nvalid ? p.a[j] - pred[j] : 0.f;
        out[j] = PLAIN ? d : e.escale * d;
      }
      if (e.O) st4(e.O, idx, nvalid, vec, out);
      if (e.red) {  // layer energy (lam/2) sum d^2 -- only the launches that report energies pay for it
        const float half = PLAIN ? 0.5f : (e.escale != 0.f ? 0.5f / e.escale : 0.f);  // out = lam d  =>  lam/2 d^2 = out^2 / (2 lam)
#pragma unroll
        for (int j = 0; j < 4; ++j) r += half * out[j] * out[j];
      }
      if (MODE < 0 && pre_act != ACT_LINEAR) {  // Form B: the operand copies carry post'(pre) . e (what travels through W)
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] *= dpre[j];
        if (e.Oh) st4(e.Oh, idx, nvalid, vec, out);
      }
      if (e.Ob) st_ob<SPLIT>(e, row, col, idx, nvalid, vec, out);
    }
    return r;
  }
  // ---- EPI_GRAD ----
  float el[4], g[4];
  const float (&z)[4] = p.a;
  if (ELSRC == 0) {
    unpack_bf16x4(p.h, el);
  } else if (ELSRC == 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) el[j] = z[j] - p.b[j];
  } else if (ELSRC == 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) el[j] = p.c[j];
  } else if (e.El) {
#pragma unroll
    for (int j = 0; j < 4; ++j) el[j] = p.b[j];
  } else if (e.Elb) {
    if (vec) unpack_bf16x4(p.h, el);
    else {
#pragma unroll
      for (int j = 0; j < 4; ++j) el[j] = p.b[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) el[j] = z[j] - p.b[j];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) g[j] = el[j] - act_bwd<FAST>(act, z[j]) * (e.alpha * acc[j]);
  if (!PLAIN && e.skip != 0.f) {
    float en[4];
    if (e.En) ld4(e.En, idx, nvalid, vec, en);
    else ld4(e.Enb, idx, nvalid, vec, en);
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] -= e.skip * en[j];
  }
  if (!PLAIN) {
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] = e.gscale * g[j] + e.ad * z[j];
  }
  // PLAIN: the weight gscale is folded into the coefficient of g below (cg); PLAIN == 2 adds the other half Gin
  if (!PLAIN && e.Gin) {
    float gi[4];
    ld4(e.Gin, idx, nvalid, vec, gi);
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] += gi[j];
  }
  const float cg = PLAIN ? e.c * e.gscale : e.c;  // (uniform: hoisted out of the patch loop)
  if (sub == GRAD_OUT) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = PLAIN == 2 ? fmaf(cg, g[j], e.c * p.b[j]) : cg * g[j];
    st4(e.O, idx, nvalid, vec, out);
    return 0.f;
  }
  if (sub == GRAD_EULER) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = PLAIN == 2 ? fmaf(-cg, g[j], fmaf(-e.c, p.b[j], z[j])) : fmaf(-cg, g[j], z[j]);
  } else if (sub == GRAD_HEUN1) {
    st4(e.G1, idx, nvalid, vec, g);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = z[j] - e.c * g[j];
  } else {  // GRAD_HEUN2
    float z0[4], g1[4];
    ld4(e.Z0, idx, nvalid, vec, z0);
    ld4(e.G1, idx, nvalid, vec, g1);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = z0[j] - 0.5f * e.c * (g1[j] + g[j]);
    if (e.red) {
      // diffrax.Heun's embedded error estimate y_err = dt (k1 - k2) / 2 = -(c/2)(g1 - g2), scaled as
      // PIDController does: y_err / (atol + rtol * max(|y0|, |y1|)); the caller takes the RMS over all entries
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nvalid) {
          const float ye = 0.5f * e.c * (g[j] - g1[j]);
          const float sc = ye / (e.tol_a + e.tol_r * fmaxf(fabsf(z0[j]), fabsf(out[j])));
          r += sc * sc;
        }
    }
  }
  st4(e.O, idx, nvalid, vec, out);
  if (e.Ob || e.Oh) {
    float h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = act_fwd<FAST>(act, out[j]);
    if (e.Ob) st_ob<SPLIT>(e, row, col, idx, nvalid, vec, h);
    if (e.Oh) st4(e.Oh, idx, nvalid, vec, h);
  }
  return r;
}

// Applies the fused epilogue to 4 horizontally adjacent accumulator values D[row, col..col+3].
template <bool FAST>
__device__ __forceinline__ float epi_apply4(const Epi& e, int row, int col, const float (&acc)[4], int nvalid, bool vec) {
  EpiPre p;
  epi_preload<-1>(e, row, col, nvalid, vec, p);
  const float nob[4] = {0.f, 0.f, 0.f, 0.f};
  return epi_finish<FAST>(e, row, col, acc, nvalid, vec, p, false, nob);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace jpcb
