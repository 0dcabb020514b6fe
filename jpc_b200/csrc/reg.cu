// reg.cu -- the weight regularisers of pc_energy_fn (jpc/_core/_energies.py:127-143), which the reference applies to the
// layers that expose `.weight` directly -- bare `nn.Linear` layers, e.g. the last layer of its hand-built example networks;
// for `make_mlp` models (Sequential layers) both are no-ops there:
//   F_reg       = wd/2 sum_l |W_l|^2 + sp/2 sum_l |W_l^T W_l - I|^2          (not divided by the batch)
//   dF_reg/dW_l = wd W_l + 2 sp W_l (W_l^T W_l - I)
// Two GEMMs per layer on the fp32 CUDA-core kernel (operands read in place through strides) + elementwise passes; this is
// parameter-side work done once per gradient evaluation, never inside the inference loop.
#include <cuda_runtime.h>

#include <vector>

#include "elementwise.cuh"
#include "host.cuh"

using namespace jpcb;

namespace {

// S <- S - I on the diagonal, acc += sum S^2  (S: [C, C])
__global__ void spectral_fix_kernel(float* __restrict__ S, int C, float* __restrict__ acc) {
  __shared__ float red_s[32];
  float s = 0.f;
  const size_t n = (size_t)C * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = S[i];
    if (i / C == i % C) { v -= 1.f; S[i] = v; }
    s += v * v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red_s[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(acc, v);
  }
}

__global__ void reg_finish_kernel(const float* acc, float wd, float sp, float* F, int accumulate) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float v = 0.5f * wd * acc[0] + 0.5f * sp * acc[1];
    *F = accumulate ? *F + v : v;
  }
}

int ew_grid_(size_t items, int per_block) {
  size_t blocks = (items + per_block - 1) / per_block;
  const size_t cap = (size_t)num_sms() * 8;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

size_t reg_bytes(int n, const int* cols, float sp_nonzero) {
  size_t b = 1024;  // two accumulators
  if (sp_nonzero != 0.f)
    for (int i = 0; i < n; ++i) b += (((size_t)cols[i] * cols[i] * sizeof(float)) + 1023) & ~(size_t)1023;
  return b;
}

}  // namespace

extern "C" {

size_t jpcb_pc_weight_reg_workspace_bytes(int n, const int* rows, const int* cols) {
  (void)rows;
  if (n <= 0 || !cols) return 0;
  return reg_bytes(n, cols, 1.f);
}

int jpcb_pc_weight_reg(int n, const float* const* W, const int* rows, const int* cols, float weight_decay, float spectral_penalty,
                       float* F_inout, int accumulate_F, float* const* gW_inout, void* workspace, size_t workspace_bytes, void* stream) {
  if (n <= 0 || !W || !rows || !cols) return fail(JPCB_EINVAL, "weight regularisers: empty / null tensor list");
  for (int i = 0; i < n; ++i)
    if (!W[i] || rows[i] <= 0 || cols[i] <= 0) return fail(JPCB_EINVAL, "weight regularisers: bad tensor %d", i);
  const size_t need = reg_bytes(n, cols, spectral_penalty);
  if (!workspace || workspace_bytes < need) return fail(JPCB_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* acc = reinterpret_cast<float*>(workspace);
  CK(cudaMemsetAsync(acc, 0, 2 * sizeof(float), st));
  char* p = reinterpret_cast<char*>(workspace) + 1024;
  for (int i = 0; i < n; ++i) {
    const int R = rows[i], C = cols[i];
    const size_t nW = (size_t)R * C;
    if (weight_decay != 0.f) {
      if (F_inout) { sumsq_kernel<<<ew_grid_(nW, 1024), 256, 0, st>>>(W[i], nW, acc); count_launch(); }
      if (gW_inout && gW_inout[i]) { axpby_kernel<<<ew_grid_(nW, 256), 256, 0, st>>>(gW_inout[i], W[i], gW_inout[i], nW, 1.f, weight_decay); count_launch(); }
    }
    if (spectral_penalty != 0.f) {
      float* S = reinterpret_cast<float*>(p);
      p += (((size_t)C * C * sizeof(float)) + 1023) & ~(size_t)1023;
      GTask t;  // S = W^T W: D[m, n] = sum_k W[k][m] W[k][n]
      t.e.mode = EPI_SCALE; t.e.M = C; t.e.N = C; t.K = R;
      t.a_p = W[i]; t.a_smn = 1; t.a_sk = C;
      t.b_p = W[i]; t.b_smn = 1; t.b_sk = C;
      t.e.alpha = 1.f; t.e.O = S;
      RET(launch_simt({t}, st));
      spectral_fix_kernel<<<ew_grid_((size_t)C * C, 1024), 256, 0, st>>>(S, C, acc + 1); count_launch();
      if (gW_inout && gW_inout[i]) {
        GTask u;  // gW += 2 sp W (W^T W - I): D[m, n] = sum_k W[m][k] S[n][k]  (S symmetric)
        u.e.mode = EPI_SCALE; u.e.M = R; u.e.N = C; u.K = C;
        u.a_p = W[i]; u.a_smn = C; u.a_sk = 1;
        u.b_p = S; u.b_smn = C; u.b_sk = 1;
        u.e.alpha = 2.f * spectral_penalty; u.e.accum = 1; u.e.O = gW_inout[i];
        RET(launch_simt({u}, st));
      }
    }
  }
  CK(cudaGetLastError());
  if (F_inout) { reg_finish_kernel<<<1, 32, 0, st>>>(acc, weight_decay, spectral_penalty, F_inout, accumulate_F); count_launch(); }
  CK(cudaGetLastError());
  return JPCB_OK;
}

}  // extern "C"
