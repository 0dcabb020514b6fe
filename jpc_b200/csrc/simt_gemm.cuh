// simt_gemm.cuh -- grouped fp32 CUDA-core GEMM with the fused PC epilogues (common.cuh).
//
// This is the fp32 (1e-5 parity) path and the narrow-layer path: arbitrary M/N/K, arbitrary
// operand strides (so W, W^T, e^T, act(z)^T are all read in place with no transposed copies) and
// the activation applied while staging an operand into shared memory.  One launch covers every
// layer of a phase (grouped), so a T-step solve of an L-layer net is 2T launches, not 2TL.
#pragma once
#include "common.cuh"

namespace jpcb {

struct Opnd {
  const float* p;
  long long s_mn;  // element stride along M (A) or N (B)
  long long s_k;   // element stride along K
  int act;         // activation applied on load (ACT_LINEAR = none)
  int pad_;
};

struct SimtTask {
  Opnd A, B;
  int K;
  int tile_begin;  // first linear tile id of this task inside the grouped launch
  int tiles_n;     // number of tiles along N
  int pad_;
  Epi e;
};

constexpr int SIMT_MAX_TASKS = 64;
struct SimtGroup {
  int ntasks;
  int total_tiles;
  SimtTask t[SIMT_MAX_TASKS];
};

template <int BM, int BN, int BK, int RM, int RN>
__global__ void __launch_bounds__((BM / (4 * RM)) * (BN / (4 * RN)))
simt_grouped_gemm(const __grid_constant__ SimtGroup g) {
  constexpr int TY = BM / (4 * RM), TX = BN / (4 * RN), NT = TX * TY;
  constexpr int LA = (BM * BK + NT - 1) / NT, LB = (BN * BK + NT - 1) / NT;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  __shared__ float red_s[32];

  const int tid = threadIdx.x;
  int ti = 0;
  while (ti + 1 < g.ntasks && (int)blockIdx.x >= g.t[ti + 1].tile_begin) ++ti;
  const SimtTask& t = g.t[ti];
  const Epi& e = t.e;
  const int local = blockIdx.x - t.tile_begin;
  const int m0 = (local / t.tiles_n) * BM, n0 = (local % t.tiles_n) * BN;
  const int M = e.M, N = e.N, K = t.K;
  const float* __restrict__ Ap = t.A.p;
  const float* __restrict__ Bp = t.B.p;
  const long long a_smn = t.A.s_mn, a_sk = t.A.s_k, b_smn = t.B.s_mn, b_sk = t.B.s_k;
  const int a_act = t.A.act, b_act = t.B.act;
  const bool a_kc = (a_sk == 1), b_kc = (b_sk == 1);  // K-contiguous operand?

  float ra[LA], rb[LB];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int idx = tid + i * NT;
      int m = a_kc ? idx / BK : idx % BM, k = a_kc ? idx % BK : idx / BM;
      float v = 0.f;
      if (idx < BM * BK && m0 + m < M && k0 + k < K) {
        v = Ap[(long long)(m0 + m) * a_smn + (long long)(k0 + k) * a_sk];
        if (a_act != ACT_LINEAR) v = act_fwd<false>(a_act, v);
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int idx = tid + i * NT;
      int n = b_kc ? idx / BK : idx % BN, k = b_kc ? idx % BK : idx / BN;
      float v = 0.f;
      if (idx < BN * BK && n0 + n < N && k0 + k < K) {
        v = Bp[(long long)(n0 + n) * b_smn + (long long)(k0 + k) * b_sk];
        if (b_act != ACT_LINEAR) v = act_fwd<false>(b_act, v);
      }
      rb[i] = v;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int idx = tid + i * NT;
      if (idx < BM * BK) {
        int m = a_kc ? idx / BK : idx % BM, k = a_kc ? idx % BK : idx / BM;
        As[buf][k][m] = ra[i];
      }
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int idx = tid + i * NT;
      if (idx < BN * BK) {
        int n = b_kc ? idx / BK : idx % BN, k = b_kc ? idx % BK : idx / BN;
        Bs[buf][k][n] = rb[i];
      }
    }
  };

  const int ty = tid / TX, tx = tid % TX;
  float acc[4 * RM][4 * RN];
#pragma unroll
  for (int i = 0; i < 4 * RM; ++i)
#pragma unroll
    for (int j = 0; j < 4 * RN; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  if (nk > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload((kb + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4 * RM], b[4 * RN];
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4 + r * (BM / RM)]);
        a[4 * r] = v.x; a[4 * r + 1] = v.y; a[4 * r + 2] = v.z; a[4 * r + 3] = v.w;
      }
#pragma unroll
      for (int r = 0; r < RN; ++r) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4 + r * (BN / RN)]);
        b[4 * r] = v.x; b[4 * r + 1] = v.y; b[4 * r + 2] = v.z; b[4 * r + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < 4 * RM; ++i)
#pragma unroll
        for (int j = 0; j < 4 * RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

  // fused epilogue
  const bool nvec = (N % 4) == 0;
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 4 * RM; ++i) {
    const int row = m0 + ty * 4 + (i / 4) * (BM / RM) + (i % 4);
    if (row >= M) continue;
#pragma unroll
    for (int jr = 0; jr < RN; ++jr) {
      const int col = n0 + tx * 4 + jr * (BN / RN);
      const int nvalid = min(4, N - col);
      if (nvalid <= 0) continue;
      float a4[4] = {acc[i][4 * jr], acc[i][4 * jr + 1], acc[i][4 * jr + 2], acc[i][4 * jr + 3]};
      r += epi_apply4<false>(e, row, col, a4, nvalid, nvec && nvalid == 4);
    }
  }
  if (e.red != nullptr) {
    r = warp_sum(r);
    if ((tid & 31) == 0) red_s[tid >> 5] = r;
    __syncthreads();
    if (tid < 32) {
      float v = tid < (NT + 31) / 32 ? red_s[tid] : 0.f;
      v = warp_sum(v);
      if (tid == 0) atomicAdd(e.red, v);
    }
  }
}

}  // namespace jpcb
