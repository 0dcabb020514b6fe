// optim.cu -- the parameter update that closes the train step (jpc/_train.py:234-242:
// `optim.update(updates=param_grads, ...)` then `eqx.apply_updates`): optax.sgd / optax.adam over every
// parameter tensor of the model in ONE pass per tensor group.  HBM-bound elementwise work: each element
// of p, g (and the two Adam moments) is read once and each result written once, 16 bytes per thread per
// access, grid sized to the SM count (grid-stride), no intermediate "updates" tensor.
#include <cmath>

#include "host.cuh"

namespace jpcb {
namespace {

constexpr int OPT_MAX = 32;  // tensors per launch (pointer table lives in kernel-parameter space)

struct OptTable {
  const float* p[OPT_MAX];
  const float* g[OPT_MAX];
  const float* m[OPT_MAX];
  const float* v[OPT_MAX];
  float* po[OPT_MAX];
  float* mo[OPT_MAX];
  float* vo[OPT_MAX];
  long long start[OPT_MAX + 1];  // prefix sum of ceil(numel / 4): the unit of work is one 4-float vector
  long long numel[OPT_MAX];
  int n;
};

template <bool ADAM>
__device__ __forceinline__ void opt_elem(float p, float g, float m, float v, float lr, float b1, float b2, float eps,
                                         float omb1, float omb2, float c1, float c2, float& po, float& mo, float& vo) {
  if (ADAM) {
    // optax.scale_by_adam: mu = b1 mu + (1-b1) g; nu = b2 nu + (1-b2) g^2; bias correction by 1 - b^t (c1, c2);
    // update = -lr * mu_hat / (sqrt(nu_hat) + eps)   (eps_root = 0)
    mo = b1 * m + omb1 * g;
    vo = b2 * v + omb2 * g * g;
    po = p - lr * ((mo / c1) / (sqrtf(vo / c2) + eps));
  } else {
    po = p - lr * g;  // optax.sgd: updates = -lr g
  }
}

template <bool ADAM>
__global__ void __launch_bounds__(256) optim_kernel(const __grid_constant__ OptTable T, float lr, float b1, float b2,
                                                    float eps, float omb1, float omb2, float c1, float c2) {
  const long long total = T.start[T.n];
  int t = 0;
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < total; u += (long long)gridDim.x * blockDim.x) {
    while (u >= T.start[t + 1]) ++t;  // units are visited in increasing order: the cursor only moves forward
    const long long i = (u - T.start[t]) * 4;
    const long long left = T.numel[t] - i;
    const bool vec = left >= 4 && ((reinterpret_cast<uintptr_t>(T.p[t]) | reinterpret_cast<uintptr_t>(T.g[t]) |
                                    reinterpret_cast<uintptr_t>(T.po[t]) |
                                    (ADAM ? (reinterpret_cast<uintptr_t>(T.m[t]) | reinterpret_cast<uintptr_t>(T.v[t]) |
                                             reinterpret_cast<uintptr_t>(T.mo[t]) | reinterpret_cast<uintptr_t>(T.vo[t]))
                                          : 0)) & 15) == 0;
    if (vec) {
      const float4 p = __ldcs(reinterpret_cast<const float4*>(T.p[t] + i));
      const float4 g = __ldcs(reinterpret_cast<const float4*>(T.g[t] + i));
      float4 m = make_float4(0, 0, 0, 0), v = m, po, mo = m, vo = m;
      if (ADAM) {
        m = __ldcs(reinterpret_cast<const float4*>(T.m[t] + i));
        v = __ldcs(reinterpret_cast<const float4*>(T.v[t] + i));
      }
      opt_elem<ADAM>(p.x, g.x, m.x, v.x, lr, b1, b2, eps, omb1, omb2, c1, c2, po.x, mo.x, vo.x);
      opt_elem<ADAM>(p.y, g.y, m.y, v.y, lr, b1, b2, eps, omb1, omb2, c1, c2, po.y, mo.y, vo.y);
      opt_elem<ADAM>(p.z, g.z, m.z, v.z, lr, b1, b2, eps, omb1, omb2, c1, c2, po.z, mo.z, vo.z);
      opt_elem<ADAM>(p.w, g.w, m.w, v.w, lr, b1, b2, eps, omb1, omb2, c1, c2, po.w, mo.w, vo.w);
      *reinterpret_cast<float4*>(T.po[t] + i) = po;
      if (ADAM) {
        *reinterpret_cast<float4*>(T.mo[t] + i) = mo;
        *reinterpret_cast<float4*>(T.vo[t] + i) = vo;
      }
    } else {
      for (int j = 0; j < 4 && j < left; ++j) {
        float po, mo = 0.f, vo = 0.f;
        opt_elem<ADAM>(T.p[t][i + j], T.g[t][i + j], ADAM ? T.m[t][i + j] : 0.f, ADAM ? T.v[t][i + j] : 0.f, lr, b1, b2,
                       eps, omb1, omb2, c1, c2, po, mo, vo);
        T.po[t][i + j] = po;
        if (ADAM) {
          T.mo[t][i + j] = mo;
          T.vo[t][i + j] = vo;
        }
      }
    }
  }
}

template <bool ADAM>
int run(int n, const long long* numel, const float* const* p, const float* const* g, const float* const* m,
        const float* const* v, double lr, double b1, double b2, double eps, int step, float* const* po, float* const* mo,
        float* const* vo, cudaStream_t st) {
  if (n < 0 || (n > 0 && (!numel || !p || !g || !po))) return fail(JPCB_EINVAL, "optimiser: null pointer table");
  if (ADAM && n > 0 && (!m || !v || !mo || !vo)) return fail(JPCB_EINVAL, "adam: null moment tables");
  if (ADAM && step < 1) return fail(JPCB_EINVAL, "adam: step=%d (counts from 1)", step);
  // hyper-parameter arithmetic in double on the host, rounded once to fp32 -- as optax does with Python floats
  // (1 - b2 is fp32(0.001), not 1 - fp32(0.999))
  const float c1 = ADAM ? (float)(1.0 - pow(b1, (double)step)) : 1.f, c2 = ADAM ? (float)(1.0 - pow(b2, (double)step)) : 1.f;
  const float omb1 = (float)(1.0 - b1), omb2 = (float)(1.0 - b2);
  for (int base = 0; base < n; base += OPT_MAX) {
    OptTable T;
    memset(&T, 0, sizeof(T));
    T.n = n - base < OPT_MAX ? n - base : OPT_MAX;
    for (int i = 0; i < T.n; ++i) {
      const int k = base + i;
      if (numel[k] < 0 || (numel[k] > 0 && (!p[k] || !g[k] || !po[k]))) return fail(JPCB_EINVAL, "optimiser: tensor %d", k);
      T.p[i] = p[k], T.g[i] = g[k], T.po[i] = po[k], T.numel[i] = numel[k];
      if (ADAM) T.m[i] = m[k], T.v[i] = v[k], T.mo[i] = mo[k], T.vo[i] = vo[k];
      T.start[i + 1] = T.start[i] + (numel[k] + 3) / 4;
    }
    const long long units = T.start[T.n];
    if (units == 0) continue;
    long long blocks = (units + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    optim_kernel<ADAM><<<(unsigned)blocks, 256, 0, st>>>(T, (float)lr, (float)b1, (float)b2, (float)eps, omb1, omb2, c1, c2);
    CK(cudaGetLastError());
    count_launch();
  }
  return JPCB_OK;
}

}  // namespace
}  // namespace jpcb

extern "C" int jpcb_optim_sgd(int n_tensors, const long long* numel, const float* const* params, const float* const* grads,
                              double learning_rate, float* const* params_out, void* stream) {
  return jpcb::run<false>(n_tensors, numel, params, grads, nullptr, nullptr, learning_rate, 0.0, 0.0, 0.0, 1, params_out,
                          nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int jpcb_optim_adam(int n_tensors, const long long* numel, const float* const* params, const float* const* grads,
                               const float* const* mu, const float* const* nu, double learning_rate, double b1, double b2,
                               double eps, int step, float* const* params_out, float* const* mu_out, float* const* nu_out,
                               void* stream) {
  return jpcb::run<true>(n_tensors, numel, params, grads, mu, nu, learning_rate, b1, b2, eps, step, params_out, mu_out,
                         nu_out, (cudaStream_t)stream);
}
