// warp_gemm.cuh -- the narrow-layer fp32 path: grouped "warp-level dot product" GEMM with the fused
// PC epilogues (common.cuh).  D[M,N] = A[M,K] * B[N,K]^T with BOTH operands K-contiguous fp32.
//
// When a phase offers fewer output tiles than the chip has SMs (configs 1-3: batch 64-256, widths
// 10-784) a tiled GEMM leaves most of the machine idle and is bound by the latency of its own k-loop.
// Here every WARP owns a 4 x 4 patch of outputs and the 32 lanes split K: each lane streams 16-byte
// vectors of 4 A rows and 4 B rows (fully coalesced 512-byte row segments per warp instruction),
// accumulates 16 partial dot products, and a 16-shuffle transpose-reduction leaves every sum in one
// lane.  A 64 x 300 x 300 layer becomes 1 200 independent warps -- 8+ resident warps on every SM --
// and the whole phase is one short launch over all layers (grouped), bound by L2 bandwidth.
// Operand preparation the host guarantees: activations already applied (fp32 phi(z) copies kept
// by the epilogues), transposed fp32 weight copies for the W^T GEMMs.
#pragma once
#include "common.cuh"
#include "simt_gemm.cuh"  // SimtTask / SimtGroup

namespace jpcb {

constexpr int WG_NT = 4;     // output columns per warp (one epi_apply4 group per row)
constexpr int WG_WARPS = 8;  // warps per CTA

__device__ __forceinline__ float4 wg_ld4(const float* __restrict__ p, int k, int K, bool vec) {
  if (vec) return *reinterpret_cast<const float4*>(p + k);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < K) v.x = p[k];
  if (k + 1 < K) v.y = p[k + 1];
  if (k + 2 < K) v.z = p[k + 2];
  if (k + 3 < K) v.w = p[k + 3];
  return v;
}

// MT = 4 or 8 output rows per warp (8 when the phase still offers enough warps: 1.5 instead of 2 bytes
// of operand traffic per multiply-add).
template <int MT>
__global__ void __launch_bounds__(WG_WARPS * 32) warp_grouped_gemm(const __grid_constant__ SimtGroup g) {
  // programmatic dependent launch: let the next kernel of the stream get resident now, then wait for
  // everything before us to be complete and visible (these launches are a chain of short dependent kernels)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31;
  const int wt = (int)blockIdx.x * WG_WARPS + (threadIdx.x >> 5);  // this warp's patch in the flattened list
  if (wt >= g.total_tiles) return;
  int ti = 0;
  while (ti + 1 < g.ntasks && wt >= g.t[ti + 1].tile_begin) ++ti;
  const SimtTask& t = g.t[ti];
  const Epi& e = t.e;
  const int local = wt - t.tile_begin;
  const int m0 = (local / t.tiles_n) * MT, n0 = (local % t.tiles_n) * WG_NT;
  const int M = e.M, N = e.N, K = t.K;
  const float* __restrict__ Ap = t.A.p;
  const float* __restrict__ Bp = t.B.p;
  const long long a_smn = t.A.s_mn, b_smn = t.B.s_mn;
  // 16-byte vector loads need K % 4 == 0 and rows starting on 16-byte boundaries
  const bool vec = (K % 4 == 0) && (a_smn % 4 == 0) && (b_smn % 4 == 0) && ((reinterpret_cast<uintptr_t>(Ap) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(Bp) & 15) == 0);
  const float* ar[MT];
  const float* br[WG_NT];
#pragma unroll
  for (int i = 0; i < MT; ++i) ar[i] = Ap + (long long)min(m0 + i, M - 1) * a_smn;  // clamped rows: results ignored
#pragma unroll
  for (int j = 0; j < WG_NT; ++j) br[j] = Bp + (long long)min(n0 + j, N - 1) * b_smn;

  float acc[MT * WG_NT];
#pragma unroll
  for (int i = 0; i < MT * WG_NT; ++i) acc[i] = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    float4 a[MT], b[WG_NT];
#pragma unroll
    for (int i = 0; i < MT; ++i) a[i] = wg_ld4(ar[i], k, K, vec);
#pragma unroll
    for (int j = 0; j < WG_NT; ++j) b[j] = wg_ld4(br[j], k, K, vec);
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < WG_NT; ++j) {
        float s = acc[i * WG_NT + j];
        s = fmaf(a[i].x, b[j].x, s);
        s = fmaf(a[i].y, b[j].y, s);
        s = fmaf(a[i].z, b[j].z, s);
        s = fmaf(a[i].w, b[j].w, s);
        acc[i * WG_NT + j] = s;
      }
  }
  // Transpose-reduction over the 32 lanes: at the step with lane-bit mask `o` each lane keeps half of its
  // values and adds the partner's copy of them, so 16 values need 8+4+2+1 shuffles (+1 to fold the last
  // pair) and 32 values 16+8+4+2+1.  Afterwards value index v = i * 4 + j sits in lane 2v (MT = 4; also in
  // 2v+1) or in lane v (MT = 8): the four columns of output row i are in the four lanes of "quad" i.
  constexpr int LAST = MT * WG_NT >= 32 ? 1 : 2;
#pragma unroll
  for (int o = 16, n = MT * WG_NT; o >= LAST; o >>= 1, n >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int v = 0; v < n / 2; ++v) {
      const float keep = upper ? acc[v + n / 2] : acc[v];
      const float send = upper ? acc[v] : acc[v + n / 2];
      acc[v] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  float s = acc[0];
  if (MT * WG_NT < 32) s += __shfl_xor_sync(0xffffffffu, s, 1);
  constexpr int LPV = MT * WG_NT >= 32 ? 1 : 2;  // lanes per value
  // lane (4 * LPV * i) gathers row i's four sums
  float row[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) row[j] = __shfl_sync(0xffffffffu, s, ((lane / (4 * LPV)) * 4 + j) * LPV);
  const int i_row = lane / (4 * LPV);
  float r = 0.f;
  if ((lane % (4 * LPV)) == 0 && m0 + i_row < M) {
    const int nvalid = min(4, N - n0);
    r = epi_apply4<false>(e, m0 + i_row, n0, row, nvalid, (N % 4) == 0 && nvalid == 4);
  }
  if (e.red != nullptr) {
    r = warp_sum(r);
    if (lane == 0 && r != 0.f) atomicAdd(e.red, r);
  }
}

// ---------------------------------------------------------------------------------------------
// Wider variant for phases that offer several warps per SM even at 128 outputs per warp (configs 2, 3):
// a warp owns an 8 x 16 tile as four 8 x 4 patches, one per group of 8 lanes, and only those 8 lanes split K.
// Versus the 32-lane split: 4x the multiply-adds per warp for the same fixed costs, a 3-level (28-shuffle)
// transpose-reduction that leaves one finished output ROW SEGMENT (4 columns) in EVERY lane -- so the fused
// epilogue runs on all 32 lanes instead of 8 -- and the four groups read the same A rows (one request per
// warp instruction).  A group's 8 lanes read 128 contiguous bytes of a row per instruction.
// ---------------------------------------------------------------------------------------------
constexpr int WG8_MT = 8, WG8_NT = 16, WG8_KL = 8;

__global__ void __launch_bounds__(WG_WARPS * 32) warp8_grouped_gemm(const __grid_constant__ SimtGroup g) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31;
  const int kl = lane & (WG8_KL - 1), sub = lane >> 3;
  const int wt = (int)blockIdx.x * WG_WARPS + (threadIdx.x >> 5);
  if (wt >= g.total_tiles) return;
  int ti = 0;
  while (ti + 1 < g.ntasks && wt >= g.t[ti + 1].tile_begin) ++ti;
  const SimtTask& t = g.t[ti];
  const Epi& e = t.e;
  const int local = wt - t.tile_begin;
  const int m0 = (local / t.tiles_n) * WG8_MT, n0 = (local % t.tiles_n) * WG8_NT + sub * 4;
  const int M = e.M, N = e.N, K = t.K;
  const float* __restrict__ Ap = t.A.p;
  const float* __restrict__ Bp = t.B.p;
  const long long a_smn = t.A.s_mn, b_smn = t.B.s_mn;
  const bool vec = (K % 4 == 0) && (a_smn % 4 == 0) && (b_smn % 4 == 0) && ((reinterpret_cast<uintptr_t>(Ap) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(Bp) & 15) == 0);
  // row starts as 32-bit element offsets (every operand of this path is far below 2^31 elements; checked on the host)
  uint32_t ao[WG8_MT], bo[4];
#pragma unroll
  for (int i = 0; i < WG8_MT; ++i) ao[i] = (uint32_t)min(m0 + i, M - 1) * (uint32_t)a_smn;  // clamped rows: results ignored
#pragma unroll
  for (int j = 0; j < 4; ++j) bo[j] = (uint32_t)min(n0 + j, N - 1) * (uint32_t)b_smn;

  float acc[WG8_MT * 4];
#pragma unroll
  for (int i = 0; i < WG8_MT * 4; ++i) acc[i] = 0.f;
  for (int k = kl * 4; k < K; k += WG8_KL * 4) {
    float4 a[WG8_MT], b[4];
#pragma unroll
    for (int i = 0; i < WG8_MT; ++i) a[i] = wg_ld4(Ap + ao[i], k, K, vec);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = wg_ld4(Bp + bo[j], k, K, vec);
#pragma unroll
    for (int i = 0; i < WG8_MT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float s = acc[i * 4 + j];
        s = fmaf(a[i].x, b[j].x, s);
        s = fmaf(a[i].y, b[j].y, s);
        s = fmaf(a[i].z, b[j].z, s);
        s = fmaf(a[i].w, b[j].w, s);
        acc[i * 4 + j] = s;
      }
  }
  // transpose-reduction over the group's 8 lanes: 32 -> 16 -> 8 -> 4 values; lane bit 2, 1, 0 = row bit 2, 1, 0,
  // so lane kl ends with the four columns of row m0 + kl in acc[0..3]
#pragma unroll
  for (int o = 4, n = WG8_MT * 4; o >= 1; o >>= 1, n >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int v = 0; v < n / 2; ++v) {
      const float keep = upper ? acc[v + n / 2] : acc[v];
      const float send = upper ? acc[v] : acc[v + n / 2];
      acc[v] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  const float row[4] = {acc[0], acc[1], acc[2], acc[3]};
  float r = 0.f;
  const int nvalid = min(4, N - n0);
  if (m0 + kl < M && nvalid > 0) r = epi_apply4<false>(e, m0 + kl, n0, row, nvalid, (N % 4) == 0 && nvalid == 4);
  if (e.red != nullptr) {
    r = warp_sum(r);
    if (lane == 0 && r != 0.f) atomicAdd(e.red, r);
  }
}

}  // namespace jpcb
