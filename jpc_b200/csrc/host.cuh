// host.cuh -- host-side vocabulary shared by the translation units of libjpc_b200.so (api.cu: C-ABI,
// descriptor validation, phase schedule, CUDA-core launches; tc_gemm.cu: tensor-core launches;
// bpc.cu: bidirectional-PC schedule): error reporting, the launch counter, and GTask, the
// "one GEMM + fused epilogue" work item both kernel families are driven by.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <functional>
#include <vector>

#include "../../include/jpc_b200.h"
#include "common.cuh"

namespace jpcb {

// sets the thread-local message returned by jpcb_last_error() and returns `code`
int fail(int code, const char* fmt, ...);
// one more kernel enqueued by this library (jpcb_launch_count)
void count_launch(int n = 1);
int num_sms();
// jpcb_set_reproducible: energies / losses by fixed-order sums instead of atomics
bool reproducible();
// experiment switches from the environment variable JPCB_TC_FLAGS (bit 1: 1-CTA tensor-core kernel)
int tc_flags();

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess) return ::jpcb::fail(JPCB_ECUDA, "%s -> %s", #call, cudaGetErrorString(_e)); \
  } while (0)
#define RET(call)                 \
  do {                            \
    int _r = (call);              \
    if (_r != JPCB_OK) return _r; \
  } while (0)

// D[M,N] = A[M,K] * B[N,K]^T followed by the fused epilogue `e`.  fp32 operands are described by
// strides (so W, W^T, e^T, act(z)^T are read in place by the CUDA-core kernel); bf16 operands
// (tensor-core kernel) are dense K-major copies.
struct GTask {
  Epi e;
  int K = 0;
  const float* a_p = nullptr; long long a_smn = 0, a_sk = 0; int a_act = ACT_LINEAR;
  const float* b_p = nullptr; long long b_smn = 0, b_sk = 0; int b_act = ACT_LINEAR;
  const __nv_bfloat16* a_b = nullptr;  // [M, K] K-major   (mn_major: [K, M] row-major)
  const __nv_bfloat16* b_b = nullptr;  // [N, K] K-major   (mn_major: [K, N] row-major)
  bool mn_major = false;               // tensor-core path: D = a_b^T b_b, operands read MN-major (no transposed copies)
  // fp32-parity tensor-core path: a_b / b_b are 3-way split copies [rows, 3 P] (elementwise.cuh: split3_kernel) and the
  // GEMM runs six passes over K (hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi).  a_ld / b_ld: row pitch (3 P) of the
  // copies; a_part / b_part: width P of one part (of the K dimension, or of the M / N dimension when mn_major).
  int split = 0, a_ld = 0, b_ld = 0, a_part = 0, b_part = 0;
  GTask() { memset(&e, 0, sizeof(e)); e.gscale = 1.f; e.escale = 1.f; }
};

// one grouped launch (per <= 16 / 64 tasks) of the tcgen05 / CUDA-core kernel over `tasks`
int launch_tc(const std::vector<GTask>& tasks, cudaStream_t st);
int launch_simt(const std::vector<GTask>& tasks, cudaStream_t st);
// the whole fixed-step Euler solve as one launch (tc_staged.cuh); JPCB_SOLVE_NA: not applicable, enqueue phase by phase
constexpr int JPCB_SOLVE_NA = -1;
int launch_tc_solve(const std::vector<GTask>& errs, const std::vector<GTask>& grads, int nsteps, float c_last, bool wave,
                    int* counters, size_t counter_ints, cudaStream_t st);


// ---- launch plans (plan.cu): a call whose key (entry point, descriptor, every pointer argument, workspace, stream, device)
// has been seen `min_calls` times is captured into a CUDA graph once and replayed from then on ----
struct PlanKey {
  std::vector<uint64_t> w;
  explicit PlanKey(uint64_t entry_id) { w.reserve(640); w.push_back(entry_id); }
  void word(uint64_t v) { w.push_back(v); }
  void ptr(const void* p) { w.push_back(reinterpret_cast<uint64_t>(p)); }
  void bytes(const void* p, size_t n);
  template <typename T>
  void list(T* const* l, int n) {  // host array of device pointers (nullable)
    w.push_back(l ? (uint64_t)n : ~0ull);
    if (l) for (int i = 0; i < n; ++i) w.push_back(reinterpret_cast<uint64_t>(l[i]));
  }
};
// `body(stream)` enqueues the call's kernel sequence on the stream it is given
int plan_run(PlanKey& key, void* stream, const std::function<int(void*)>& body);

}  // namespace jpcb
