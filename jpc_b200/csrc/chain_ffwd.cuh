// chain_ffwd.cuh -- the feed-forward initialisation of a DEEP, NARROW network as one launch (fp32 path).
//
// init_activities_with_ffwd (jpc/_core/_init.py:68-80) is a chain of L dependent layer maps
//   z_{l+1} = a_l post_l(W_l act_l(z_l) + b_l) (+ z_l)
// and on the narrow path every one of them used to be its own launch: 128 dependent launches of a few microseconds each
// for BASELINE config 3 (width 128, depth 128, batch 128) -- 0.45 ms of its 1.04 ms train step for 4 MFLOP per layer.
// Samples are independent through the whole chain, so here a CTA OWNS `R` batch rows and walks all L layers for them:
// the only synchronisation is the block barrier between a layer's outputs and the next layer's inputs, and those inputs
// (raw z for the skip term, act(z) for the GEMM) never leave shared memory: the global stores of z, P1 and the operand
// copies are off the dependency chain.  Per layer the CTA's 16 warps sweep the output columns four or eight at a time with
// the lanes splitting K (16-byte coalesced reads of the K-contiguous weight rows, as warp_gemm.cuh does); every CTA streams
// every weight from L2, so this pays when the weights are small and the chain is long -- the host decides (api.cu).
// (First version staged every layer's input from global memory: 4.9 us per 128 x 128 layer, slower than the launches it
// replaced; r02 launch lists under profiles/.)
#pragma once
#include "common.cuh"

namespace jpcb {

constexpr int CF_THREADS = 512;
constexpr int CF_MAX_LAYERS = 256;

struct CfLayer {
  const float* W;     // [N, K] row-major (K contiguous)
  const float* bias;  // [N] or null
  float* O;           // z_{l+1} [B, N]
  float* O2;          // optional second copy (layer 0: the hoisted prediction P1) or null
  float* Oh;          // optional act_next(z_{l+1}) [B, N]: the next phase's fp32 operand copy, or null
  int K, N;
  int act_in, act_post, act_next;
  float alpha, skip;
};

struct CfArgs {
  int L, B;
  const float* x;    // [B, K_0]
  const float* Y;    // mse loss target [B, N_{L-1}] or null
  float* loss;       // accumulator (sum of 1/2 (y - out)^2 over the batch) or null
  CfLayer l[CF_MAX_LAYERS];
};

// One sweep of a layer's output columns by one warp: NC columns at a time, lanes split K.  `hs`: act_l(u) of the CTA's rows in
// shared memory.  Returns nothing; finished outputs go to global memory and (zn / hn) to the next layer's shared-memory inputs.
template <int R, int NC>
__device__ __forceinline__ void cf_layer(const CfLayer& Lr, const float* __restrict__ zs, const float* __restrict__ hs, float* zn, float* hn,
                                         int kpad, int row0, int B, bool last, const float* Y, float& loss, int warp, int lane, int nwarps) {
  const int K = Lr.K, N = Lr.N;
  const bool vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(Lr.W) & 15) == 0);
  for (int j0 = warp * NC; j0 < N; j0 += nwarps * NC) {
    float acc[R * NC];   // value v = r * NC + c
#pragma unroll
    for (int v = 0; v < R * NC; ++v) acc[v] = 0.f;
    const float* wr[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) wr[c] = Lr.W + (size_t)min(j0 + c, N - 1) * K;  // clamped columns: results ignored
    for (int k = lane * 4; k < K; k += 128) {
      float4 wv[NC], hv[R];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (vec) wv[c] = __ldg(reinterpret_cast<const float4*>(wr[c] + k));
        else {
          wv[c].x = wr[c][k];
          wv[c].y = k + 1 < K ? wr[c][k + 1] : 0.f;
          wv[c].z = k + 2 < K ? wr[c][k + 2] : 0.f;
          wv[c].w = k + 3 < K ? wr[c][k + 3] : 0.f;
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) hv[r] = *reinterpret_cast<const float4*>(hs + r * kpad + k);
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < NC; ++c)
          acc[r * NC + c] = fmaf(hv[r].x, wv[c].x, fmaf(hv[r].y, wv[c].y, fmaf(hv[r].z, wv[c].z, fmaf(hv[r].w, wv[c].w, acc[r * NC + c]))));
    }
    // Transpose-reduction over the 32 lanes (as in warp_gemm.cuh): at the step with lane-bit mask `o` every lane keeps half
    // of its values and adds the partner's copy of them -- 16+8+4+2+1 shuffles for 32 values (8+4+2+1, +1 to fold the last
    // pair, for 16) instead of 5 per value.  Afterwards value v sits in lane v (32 values) or in lanes 2v, 2v+1 (16 values).
    // (The first version reduced every value with its own butterfly: 160 shuffles per sweep, and ncu showed the kernel
    //  waiting on exactly those -- short-scoreboard stalls, 4.1 us per 128 x 128 layer.)
    constexpr int NV = R * NC, LAST = NV >= 32 ? 1 : 2, LPV = NV >= 32 ? 1 : 2;
    static_assert(NV == 16 || NV == 32, "chain_ffwd: 16 or 32 values per warp sweep");
#pragma unroll
    for (int o = 16, n = NV; o >= LAST; o >>= 1, n >>= 1) {
      const bool upper = (lane & o) != 0;
#pragma unroll
      for (int v = 0; v < n / 2; ++v) {
        const float keep = upper ? acc[v + n / 2] : acc[v];
        const float send = upper ? acc[v] : acc[v + n / 2];
        acc[v] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    float mine = acc[0];
    if (NV < 32) mine += __shfl_xor_sync(0xffffffffu, mine, 1);
    if (lane % LPV == 0) {
      const int v = lane / LPV, r = v / NC, j = j0 + v % NC, row = row0 + r;
      if (j < N) {
        float pre = mine + (Lr.bias ? Lr.bias[j] : 0.f);
        if (Lr.act_post != ACT_LINEAR) pre = act_fwd<false>(Lr.act_post, pre);
        float z = Lr.alpha * pre;
        if (Lr.skip != 0.f) z += Lr.skip * zs[r * kpad + j];   // (identity skip: K == N, validated on the host)
        const float h = act_fwd<false>(Lr.act_next, z);
        if (!last) { zn[r * kpad + j] = z; hn[r * kpad + j] = h; }   // the next layer's inputs never leave the SM
        if (row < B) {
          const size_t o = (size_t)row * N + j;
          Lr.O[o] = z;
          if (Lr.O2) Lr.O2[o] = z;
          if (Lr.Oh) Lr.Oh[o] = h;
          if (last && Y) { const float d = Y[o] - z; loss += 0.5f * d * d; }
        }
      }
    }
  }
}

template <int R>
__global__ void __launch_bounds__(CF_THREADS) chain_ffwd_kernel(const __grid_constant__ CfArgs a, int kpad) {
  // [2][2][R][kpad]: raw z and act(z) of this CTA's rows, double-buffered over layers; kpad % 4 == 0, kpad >= max dim + 4
  extern __shared__ __align__(16) float cf_sm[];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = CF_THREADS / 32;
  const int row0 = (int)blockIdx.x * R;
  auto zbuf = [&](int b) { return cf_sm + (size_t)(2 * b) * R * kpad; };
  auto hbuf = [&](int b) { return cf_sm + (size_t)(2 * b + 1) * R * kpad; };
  for (int idx = tid; idx < 4 * R * kpad; idx += CF_THREADS) cf_sm[idx] = 0.f;  // (ragged K reads up to 3 elements past a row)
  __syncthreads();
  {
    const int K = a.l[0].K;
    for (int idx = tid; idx < R * K; idx += CF_THREADS) {
      const int r = idx / K, k = idx - r * K;
      const float v = row0 + r < a.B ? a.x[(size_t)(row0 + r) * K + k] : 0.f;
      zbuf(0)[r * kpad + k] = v;
      hbuf(0)[r * kpad + k] = act_fwd<false>(a.l[0].act_in, v);
    }
  }
  __syncthreads();
  float loss = 0.f;
  for (int li = 0; li < a.L; ++li) {
    const CfLayer& Lr = a.l[li];
    const int cur = li & 1, nxt = cur ^ 1;
    const bool last = li == a.L - 1;
    if (Lr.N >= nwarps * 8)
      cf_layer<R, 8>(Lr, zbuf(cur), hbuf(cur), zbuf(nxt), hbuf(nxt), kpad, row0, a.B, last, a.loss ? a.Y : nullptr, loss, warp, lane, nwarps);
    else
      cf_layer<R, 4>(Lr, zbuf(cur), hbuf(cur), zbuf(nxt), hbuf(nxt), kpad, row0, a.B, last, a.loss ? a.Y : nullptr, loss, warp, lane, nwarps);
    if (!last && tid < 4 * R) {  // the next layer reads rows of N floats: keep the 3 elements past them zero (a wider layer wrote there)
      zbuf(nxt)[(tid >> 2) * kpad + Lr.N + (tid & 3)] = 0.f;
      hbuf(nxt)[(tid >> 2) * kpad + Lr.N + (tid & 3)] = 0.f;
    }
    __syncthreads();
  }
  if (a.loss) {
    loss = warp_sum(loss);
    if (lane == 0 && loss != 0.f) atomicAdd(a.loss, loss);
  }
}

}  // namespace jpcb
