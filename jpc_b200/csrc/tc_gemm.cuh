// tc_gemm.cuh -- grouped, persistent, warp-specialised tcgen05 GEMM for sm_100a with the fused PC
// epilogues of common.cuh.  D[M,N] = A[M,K] * B[N,K]^T, A and B bf16 K-major in HBM, fp32
// accumulation in TMEM.
//
//   warp 0      TMA producer : cp.async.bulk.tensor (SWIZZLE_128B boxes of 64 bf16 along K) into a
//                              4-stage shared-memory ring, completion on `full` mbarriers
//   warp 1      MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                              (128 x 256 x 16 per instruction), tcgen05.commit frees ring slots and
//                              publishes the accumulator; also owns the TMEM allocation (512 columns
//                              = two 128x256 fp32 accumulator stages, so tile i+1's MMAs overlap
//                              tile i's epilogue)
//   warps 2..9  epilogue     : tcgen05.ld the accumulator (lane = row), transpose 32x16 patches
//                              through XOR-swizzled shared memory so that all global traffic of
//                              the fused epilogue is row-contiguous 16-byte vectors, then
//                              epi_apply4 (activation, derivative, error, Euler/Heun update,
//                              bf16 operand copy for the next GEMM, energy reduction)
//
// One launch processes every layer of a phase (grouped): tiles of all tasks are flattened and
// CTA c handles tiles c, c+grid, c+2*grid, ...  (grid = min(#tiles, #SMs), 1 CTA/SM).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace jpcb {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_N = 256;
constexpr int TC_BLOCK_K = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int TC_UMMA_K = 16;
constexpr int TC_STAGES = 4;
constexpr int TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;
constexpr int TC_B_BYTES = TC_BLOCK_N * TC_BLOCK_K * 2;
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_XPOSE_BYTES = 32 * 16 * 4;  // one 32x16 fp32 patch per epilogue warp
constexpr int TC_SMEM_BYTES = 1024 /*align slack*/ + TC_STAGES * TC_STAGE_BYTES + TC_EPI_WARPS * TC_XPOSE_BYTES + 256;
constexpr int TC_TMEM_COLS = 512;

struct alignas(64) TcTask {
  CUtensorMap tmA;  // bf16 [M, K] row-major, box {64, 128}
  CUtensorMap tmB;  // bf16 [N, K] row-major, box {64, 256}
  int K;
  int tile_begin;
  int tiles_n;
  int pad_;
  Epi e;
};

constexpr int TC_MAX_TASKS = 16;
struct alignas(64) TcGroup {
  int ntasks;
  int total_tiles;
  int pad_[14];
  TcTask t[TC_MAX_TASKS];
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU (a hung box is a lost GPU lease), so after
// ~4 s of spinning the kernel traps and the host sees a launch failure instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 "version 1"): K-major operand, SWIZZLE_128B, rows of
// 128 B, 8-row groups 1024 B apart (SBO), LBO = 1 (unused for swizzled K-major).
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// Instruction descriptor for kind::f16: D=f32, A=B=bf16, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ int tc_find_task(const TcGroup& g, int tile) {
  int ti = 0;
  while (ti + 1 < g.ntasks && tile >= g.t[ti + 1].tile_begin) ++ti;
  return ti;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) tc_grouped_gemm(const __grid_constant__ TcGroup g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t xpose_base = smem_base + TC_STAGES * TC_STAGE_BYTES;
  float* xpose_ptr = reinterpret_cast<float*>(smem + TC_STAGES * TC_STAGE_BYTES);
  const uint32_t bar_base = xpose_base + TC_EPI_WARPS * TC_XPOSE_BYTES;
  // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]; then the TMEM base address
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * TC_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * TC_STAGES + 2 + s); };
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem + TC_STAGES * TC_STAGE_BYTES + TC_EPI_WARPS * TC_XPOSE_BYTES + 8 * (2 * TC_STAGES + 4));
  const uint32_t tmem_slot_addr = bar_base + 8u * (2 * TC_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), TC_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot_addr, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const int ti = tc_find_task(g, tile);
      const TcTask& t = g.t[ti];
      const int local = tile - t.tile_begin;
      const int m0 = (local / t.tiles_n) * TC_BLOCK_M, n0 = (local % t.tiles_n) * TC_BLOCK_N;
      const int nkb = (t.K + TC_BLOCK_K - 1) / TC_BLOCK_K;
      if (lane == 0) {
        tma_prefetch_desc(&t.tmA);
        tma_prefetch_desc(&t.tmB);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), TC_STAGE_BYTES);
          const uint32_t sa = smem_base + stage * TC_STAGE_BYTES;
          tma_load_2d(sa, &t.tmA, full_bar(stage), kb * TC_BLOCK_K, m0);
          tma_load_2d(sa + TC_A_BYTES, &t.tmB, full_bar(stage), kb * TC_BLOCK_K, n0);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      // keep the warp's lanes in lock-step with the elected lane's ring position
      stage = __shfl_sync(0xffffffffu, stage, 0);
      phase = __shfl_sync(0xffffffffu, phase, 0);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(TC_BLOCK_M, TC_BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const int ti = tc_find_task(g, tile);
      const int nkb = (g.t[ti].K + TC_BLOCK_K - 1) / TC_BLOCK_K;
      if (lane == 0) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * TC_BLOCK_N);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * TC_STAGE_BYTES;
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sa + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
            // advancing K inside the 128-byte swizzle row: +32 B => +2 in the (addr >> 4) field
            umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // ring slot reusable once these MMAs have read it
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(as));  // accumulator complete
      }
      stage = __shfl_sync(0xffffffffu, stage, 0);
      phase = __shfl_sync(0xffffffffu, phase, 0);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;          // 0..7
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int half = ew >> 2;         // which 128 columns of the 256-wide tile
    float* xp = xpose_ptr + ew * (TC_XPOSE_BYTES / 4);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const int ti = tc_find_task(g, tile);
      const TcTask& t = g.t[ti];
      const Epi& e = t.e;
      const int local = tile - t.tile_begin;
      const int m0 = (local / t.tiles_n) * TC_BLOCK_M, n0 = (local % t.tiles_n) * TC_BLOCK_N;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      float r = 0.f;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TC_BLOCK_N + half * 128);
#pragma unroll 1
      for (int ch = 0; ch < 8; ++ch) {
        const int cb = n0 + half * 128 + ch * 16;  // first column of this 16-wide patch
        if (cb >= e.N) break;                       // warp-uniform
        uint32_t v[16];
        tmem_ld16(t_row + (uint32_t)(ch * 16), v);
        tmem_ld_wait();
        // row `lane` of the patch -> swizzled smem (16-byte chunk k of row r sits at r*4 + (k ^ ((r>>1)&3)))
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int pos = lane * 4 + (k ^ ((lane >> 1) & 3));
          *reinterpret_cast<uint4*>(xp + pos * 4) = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
        __syncwarp();
        const int c4 = lane & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rl = (lane >> 2) + 8 * i;
          const int pos = rl * 4 + (c4 ^ ((rl >> 1) & 3));
          const float4 a = *reinterpret_cast<const float4*>(xp + pos * 4);
          const int row = m0 + q * 32 + rl;
          const int col = cb + 4 * c4;
          if (row < e.M && col < e.N) {
            const float a4[4] = {a.x, a.y, a.z, a.w};
            r += epi_apply4<true>(e, row, col, a4, 4, true);
          }
        }
        __syncwarp();
      }
      // all of this warp's TMEM reads are complete (tcgen05.wait::ld above): release the stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (e.red != nullptr) {
        r = warp_sum(r);
        if (lane == 0) atomicAdd(e.red, r);
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

}  // namespace jpcb
