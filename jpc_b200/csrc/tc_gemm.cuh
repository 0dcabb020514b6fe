// tc_gemm.cuh -- grouped, persistent, warp-specialised tcgen05 GEMM for sm_100a with the fused PC
// epilogues of common.cuh.  D[M,N] = A[M,K] * B[N,K]^T, A and B bf16 K-major in HBM, fp32
// accumulation in TMEM.
//
// The kernel is a template on CG, the tcgen05 cta_group:
//   CG = 2 (default): thread-block clusters of two CTAs (one SM pair) own a 256 x 256 output tile.
//        Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 N-rows) per
//        64-wide K block -- 32 KB per stage instead of 48 KB, 6 stages instead of 4 -- and the
//        leader CTA issues tcgen05.mma.cta_group::2 (256 x 256 x 16), which reads A and B from
//        both CTAs' shared memory and writes each CTA's 128 x 256 accumulator half into that CTA's
//        TMEM.  Per-SM shared-memory fill traffic per flop drops by a third, which is what the
//        main loop of the 1-CTA version was bound by.
//   CG = 1: one CTA per 128 x 256 tile (kept for A/B measurements: JPCB_TC_FLAGS bit 1).
//
//   warp 0      TMA producer : cp.async.bulk.tensor (SWIZZLE_128B boxes of 64 bf16 along K) into the
//                              shared-memory ring; completion (of BOTH CTAs' bytes when CG = 2) on
//                              the leader's `full` mbarriers
//   warp 1      MMA issuer   : one elected lane of the leader CTA issues tcgen05.mma; tcgen05.commit
//                              (multicast to both CTAs) frees ring slots and publishes the
//                              accumulator; also owns the TMEM allocation (512 columns = two
//                              accumulator stages, so tile i+1's MMAs overlap tile i's epilogue)
//   warps 2..9  epilogue     : per 32x16 patch: issue all global loads of the fused epilogue (one
//                              patch ahead, so two patches of loads are in flight), tcgen05.ld the
//                              accumulator (lane = row), transpose it through XOR-swizzled shared
//                              memory so that all global traffic is row-contiguous 16-byte vectors,
//                              then epi_finish (activation, derivative, error, Euler/Heun update,
//                              bf16 operand copy for the next GEMM, energy reduction), specialised
//                              at compile time per mode / activation.
//
// One launch processes every layer of a phase (grouped): tiles of all tasks are flattened and
// cluster c handles tiles c, c+#clusters, ...  (grid = min(CG * #tiles, #SMs), 1 CTA/SM).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace jpcb {

constexpr int TC_BLOCK_M = 128;  // rows per CTA (a CG = 2 cluster tile has 256)
constexpr int TC_BLOCK_N = 256;
constexpr int TC_BLOCK_K = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int TC_UMMA_K = 16;
constexpr int TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;
#ifndef JPCB_TC_EPI_WARPS
#define JPCB_TC_EPI_WARPS 8
#endif
constexpr int TC_EPI_WARPS = JPCB_TC_EPI_WARPS;  // default number of epilogue warps: 8, 12 or 16 (multiples of the 4 TMEM lane quarters)
// Epilogue-warp geometry for EW epilogue warps (a kernel template parameter: the launch-bound memory-heavy GRAD variant
// runs 12, the others 8).  The tile's 16 patches (16 columns each) are dealt to the EW / 4 column parts as evenly as
// possible: 8 warps 8+8, 12 warps 5+5+6, 16 warps 4x4.
template <int EW, int BN = 256>
struct TcEpi {
  static constexpr int NP = BN / 16;                    // 16-column patches across the tile
  static constexpr int PARTS = EW / 4;                  // column parts of the tile (one per warp of a lane quarter)
  static constexpr bool EVEN = (NP % PARTS) == 0;
  static constexpr int PATCHES = NP / PARTS;            // patches per part when the split is even
  static constexpr int THREADS = 64 + 32 * EW;          // TMA warp + MMA warp + epilogue warps
  __host__ __device__ static constexpr int part_begin(int p) { return (NP * p) / PARTS; }
};
constexpr int TC_XPOSE_BYTES = 32 * 16 * 4;  // one 32x16 fp32 patch per epilogue warp
constexpr int TC_TMEM_COLS = 512;

// BN: tile width.  256 is the default (most operand reuse per staged byte); 128 halves the tile for launches that offer
// only one or two waves of 256-wide tiles (bPC's gradient halves: 80 tile pairs on 74 CTA pairs), at a deeper ring.
template <int CG, int BN = TC_BLOCK_N>
struct TcCfg {
  static constexpr int B_ROWS = BN / CG;                   // N-rows of B staged by each CTA
  static constexpr int B_BYTES = B_ROWS * TC_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int STAGES = CG == 2 ? (BN == 256 ? 6 : 8) : 4;
  static constexpr int smem_bytes(int ew) { return 1024 /*align slack*/ + STAGES * STAGE_BYTES + ew * TC_XPOSE_BYTES + 256; }
};

struct alignas(64) TcTask {
  CUtensorMap tmA;  // bf16 [M, K] row-major, box {64, 128}        (MN-major launches: [K, M] row-major, box {64, 64})
  CUtensorMap tmB;  // bf16 [N, K] row-major, box {64, 256 / CG}   (MN-major launches: [K, N] row-major, box {64, 64})
  int K;
  int tile_begin;
  int tiles_n;
  int split;           // 0: one pass over K; 3..6: that many products over 3-way split operand copies (fp32-parity path)
  int a_part, b_part;  // split: width of one part of A / B along the concatenated dimension
  int pad_[2];
  Epi e;
};

// the six (A part, B part) products of the split path, smallest terms first: (lo,hi) (hi,lo) (mid,mid) (mid,hi) (hi,mid) (hi,hi)
__device__ __forceinline__ int tc_split_a(int pass) { return (0x052 >> (2 * pass)) & 3; }  // 2, 0, 1, 1, 0, 0 (2 bits each)
__device__ __forceinline__ int tc_split_b(int pass) { return (0x118 >> (2 * pass)) & 3; }  // 0, 2, 1, 0, 1, 0

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
// true in exactly one (elected) lane of a fully active warp; ptxas knows the region it guards is
// single-threaded, so tcgen05 / TMA instructions inside it need no per-instruction election loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMA tile load into this CTA's shared memory.  CG = 2: completion bytes are signalled on `bar`,
// a shared::cluster address that may belong to the peer (leader) CTA.
template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c_inner, int c_outer) {
  if constexpr (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c_inner), "r"(c_outer)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c_inner), "r"(c_outer)
        : "memory");
  }
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---- thread-block cluster helpers (CG = 2) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory location in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on an mbarrier given by a shared::cluster address (possibly in the peer CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
  // relaxed: the arrive only hands a TMEM stage back (its reads completed at tcgen05.wait::ld); a
  // release here would first drain all of the epilogue's global stores
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; CG = 2: M = 256 over the CTA pair, operands from both CTAs
template <int CG>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// mbarrier arrive once all previously issued MMAs have completed; CG = 2: on the barrier at the
// same offset in BOTH CTAs of the pair
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
  }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
        "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
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
// MN-major operand (the GEMM's K runs across ROWS of the source tensor: weight gradients, K = batch),
// SWIZZLE_128B: the tile is a grid of atoms of 8 K-rows x 128 B (64 bf16 along M/N), as TMA boxes of
// {64 elements, 64 rows} lay them down: 8-row groups 1024 B apart (SBO), the next 64 M/N elements in
// the next box, TC_MN_BOX_BYTES further (LBO).
constexpr int TC_MN_BOX_BYTES = 64 * TC_BLOCK_K * 2;  // one {64 x 64} bf16 box = 8 KB
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(TC_MN_BOX_BYTES >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor for kind::f16: D=f32, A=B=bf16, M x N tile; both operands K-major, or both MN-major.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, bool mn_major = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ int tc_find_task(const TcGroup& g, int tile) {
  int ti = 0;
  while (ti + 1 < g.ntasks && tile >= g.t[ti + 1].tile_begin) ++ti;
  return ti;
}

// ---------------------------------------------------------------------------------------------
// Epilogue of one warp's 32-row x 128-column share of a tile: 8 patches of 32 rows x 16 columns.
// Per patch: (1) the global loads the fused epilogue needs were issued one patch earlier (the
// first patch's even before the accumulator is complete), in the coalesced layout (4 lanes x 16 B
// = one 64-byte row segment, 8 rows per instruction; unconditional, on clamped coordinates), so
// two patches of loads are in flight per thread; (2) tcgen05.ld the accumulator patch (lane =
// row); (3) transpose it through XOR-swizzled shared memory into the same coalesced layout;
// (4) finish + store.  MODE/SUB/ACT are compile-time so that (4) is straight-line code
// (-1 = generic: read from the descriptor).
template <int MODE, int ELSRC, int PLAIN>
__device__ __forceinline__ void tc_epi_loads(const Epi& e, int row0, int cb, int lane, EpiPre (&pre)[4], float (&bias4)[4]) {
  const int c4 = lane & 3, rq = lane >> 2;
  const int colc = min(cb + 4 * c4, e.N - 4);  // N % 8 == 0 on this path
#pragma unroll
  for (int i = 0; i < 4; ++i) epi_preload<MODE, ELSRC, PLAIN>(e, min(row0 + rq + 8 * i, e.M - 1), colc, 4, true, pre[i]);
  if (MODE == EPI_FFWD || MODE == EPI_ERR || MODE < 0) {
    if (e.bias != nullptr) {
      const float4 bv = *reinterpret_cast<const float4*>(e.bias + colc);
      bias4[0] = bv.x; bias4[1] = bv.y; bias4[2] = bv.z; bias4[3] = bv.w;
    }
  }
}

template <int MODE, int SUB, int ACT, int ELSRC, int PLAIN>
__device__ __forceinline__ float tc_epi_patch(const Epi& e, int row0, int cb, int lane, uint32_t taddr, float* xp,
                                              const EpiPre (&pre)[4], const float (&bias4)[4]) {
  const int c4 = lane & 3, rq = lane >> 2;
  const int col = cb + 4 * c4;
  const bool cok = col < e.N;
  uint4* xp4 = reinterpret_cast<uint4*>(xp);
  uint32_t v[16];
  tmem_ld16(taddr, v);
  tmem_ld_wait();
  // row `lane` of the patch -> smem; 16-byte chunk k of row r sits at r*4 + (k ^ ((r >> 1) & 3))
#pragma unroll
  for (int k = 0; k < 4; ++k)
    xp4[lane * 4 + (k ^ ((lane >> 1) & 3))] = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  __syncwarp();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = rq + 8 * i;
    const float4 a = *reinterpret_cast<const float4*>(xp4 + rl * 4 + (c4 ^ ((rl >> 1) & 3)));
    const int row = row0 + rl;
    if (cok && row < e.M) {
      const float a4[4] = {a.x, a.y, a.z, a.w};
      // (fast activation approximations only in the specialised bf16 variants; the generic instantiation, which
      //  also serves the fp32-parity split path, evaluates them accurately)
      r += epi_finish<(MODE >= 0), MODE, SUB, ACT, ELSRC, PLAIN>(e, row, col, a4, 4, true, pre[i], true, bias4);
    }
  }
  __syncwarp();
  return r;
}

// Loads run TC_EPI_DEPTH - 1 patches ahead of their use (register ring of TC_EPI_DEPTH patch buffers);
// the ring index must be static, so the 8-patch loop is fully unrolled when the depth is 3.
#ifndef JPCB_TC_EPI_DEPTH
#define JPCB_TC_EPI_DEPTH 3
#endif
template <int MODE, int SUB, int ACT, int ELSRC, int PLAIN, int EW, int BN>
__device__ __forceinline__ float tc_epi_warp(const Epi& e, int row0, int col0, int lane, uint32_t t_row, float* xp, uint32_t tfull,
                                             uint32_t tphase, int np) {
  float r = 0.f;
  // (the 3-deep ring needs a fully unrolled, compile-time patch count: even splits only)
  // patches per part, rounded up when the split is uneven (the loop is fully unrolled for the static ring index and every
  // iteration is guarded by the part's own count `np`)
  using Ep = TcEpi<EW, BN>;
  constexpr int TC_EPI_PATCHES = (Ep::NP + Ep::PARTS - 1) / Ep::PARTS;
#ifndef JPCB_TC_EPI_DEPTH3_UNEVEN
#define JPCB_TC_EPI_DEPTH3_UNEVEN 0
#endif
  constexpr int DEPTH = (PLAIN == 1 && JPCB_TC_EPI_DEPTH == 3 && (Ep::EVEN || JPCB_TC_EPI_DEPTH3_UNEVEN)) ? 3 : 2;
  if constexpr (DEPTH == 3) {
    EpiPre pre[3][4];
    float bias[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    tc_epi_loads<MODE, ELSRC, PLAIN>(e, row0, col0, lane, pre[0], bias[0]);
    tc_epi_loads<MODE, ELSRC, PLAIN>(e, row0, col0 + 16, lane, pre[1], bias[1]);
    mbar_wait(tfull, tphase);  // accumulator of this tile complete
    tc_fence_after();
#pragma unroll
    for (int ch = 0; ch < TC_EPI_PATCHES; ++ch) {
      const int cb = col0 + ch * 16;  // first column of this 16-wide patch
      if ((Ep::EVEN || ch < np) && cb < e.N) {      // warp-uniform
        if (ch + 2 < TC_EPI_PATCHES && (Ep::EVEN || ch + 2 < np)) tc_epi_loads<MODE, ELSRC, PLAIN>(e, row0, cb + 32, lane, pre[(ch + 2) % 3], bias[(ch + 2) % 3]);
        r += tc_epi_patch<MODE, SUB, ACT, ELSRC, PLAIN>(e, row0, cb, lane, t_row + (uint32_t)(ch * 16), xp, pre[ch % 3], bias[ch % 3]);
      }
    }
  } else {
    EpiPre pa[4], pb[4];
    float ba[4] = {0.f, 0.f, 0.f, 0.f}, bb[4] = {0.f, 0.f, 0.f, 0.f};
    tc_epi_loads<MODE, ELSRC, PLAIN>(e, row0, col0, lane, pa, ba);
    mbar_wait(tfull, tphase);  // accumulator of this tile complete
    tc_fence_after();
#pragma unroll 1
    for (int ch = 0; ch < np; ch += 2) {   // np: this part's patch count (runtime when the split is uneven)
      const int cb = col0 + ch * 16;  // first column of this 16-wide patch
      if (cb >= e.N) break;           // warp-uniform
      const bool second = ch + 1 < np && cb + 16 < e.N;
      if (second) tc_epi_loads<MODE, ELSRC, PLAIN>(e, row0, cb + 16, lane, pb, bb);
      r += tc_epi_patch<MODE, SUB, ACT, ELSRC, PLAIN>(e, row0, cb, lane, t_row + (uint32_t)(ch * 16), xp, pa, ba);
      if (!second) break;
      if (ch + 2 < np) tc_epi_loads<MODE, ELSRC, PLAIN>(e, row0, cb + 32, lane, pa, ba);
      r += tc_epi_patch<MODE, SUB, ACT, ELSRC, PLAIN>(e, row0, cb + 16, lane, t_row + (uint32_t)(ch * 16 + 16), xp, pb, bb);
    }
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// MODE/SUB/ACT/ELSRC >= 0: every task of the launch has that epilogue mode / GRAD sub-mode /
// activation / e_l source (the host checks), compiled in; -1 = generic descriptor-driven epilogue.
// PLAIN (1 / 2): no task of the launch has a skip term, bias or decay (2: each has a partial-gradient input).
// MNMAJ: both operands are read MN-major, i.e. D = A^T B with A [K, M] and B [K, N] row-major in HBM
// (weight gradients e^T phi(z), K = batch) -- no transposed operand copies are ever materialised.
template <int CG, int MODE, int SUB, int ACT, int ELSRC, int PLAIN, bool MNMAJ = false, int EW = TC_EPI_WARPS, int BN = TC_BLOCK_N>
__global__ void __launch_bounds__(TcEpi<EW>::THREADS, 1) tc_grouped_gemm(const __grid_constant__ TcGroup g) {
  using Cfg = TcCfg<CG, BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int ACC = TC_TMEM_COLS / BN;  // fp32 accumulator stages in TMEM (512 columns): 2 at BN = 256, 4 at BN = 128
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t xpose_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  float* xpose_ptr = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  const uint32_t bar_base = xpose_base + EW * TC_XPOSE_BYTES;
  // barriers: full[STAGES], empty[STAGES], tmem_full[ACC], tmem_empty[ACC]; then the TMEM base address
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC + s); };
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem + STAGES * Cfg::STAGE_BYTES + EW * TC_XPOSE_BYTES + 8 * (2 * STAGES + 2 * ACC));
  const uint32_t tmem_slot_addr = bar_base + 8u * (2 * STAGES + 2 * ACC);
  static_assert(8 * (2 * STAGES + 2 * ACC) + 4 <= 256, "barrier block exceeds its slack in TcCfg::smem_bytes");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;  // 0 = leader CTA of the pair
  const int unit = (int)blockIdx.x / CG, nunits = (int)gridDim.x / CG;  // tile-owning CTA (pair) and their number

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);   // the leader's arrive.expect_tx (covers both CTAs' bytes)
      mbar_init(empty_bar(s), 1);  // tcgen05.commit (multicast to both CTAs when CG = 2)
    }
    for (int s = 0; s < ACC; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), EW * CG);  // the epilogue warps of both CTAs arrive on the leader's
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<CG>(tmem_slot_addr, TC_TMEM_COLS);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anyone signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation) touched no global
  // memory and may overlap the tail of the previous kernel in the stream; from here on we read what it
  // wrote.  Our own dependents may be scheduled as soon as SMs free up (they wait the same way).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = unit; tile < g.total_tiles; tile += nunits) {
      const int ti = tc_find_task(g, tile);
      const TcTask& t = g.t[ti];
      const int local = tile - t.tile_begin;
      const int m0 = (local / t.tiles_n) * (TC_BLOCK_M * CG) + (int)rank * TC_BLOCK_M;
      const int n0 = (local % t.tiles_n) * BN + (int)rank * Cfg::B_ROWS;
      const int nkb = (t.K + TC_BLOCK_K - 1) / TC_BLOCK_K;
      // t.split = number of split products to run (6 = all; 4 drops the two hi*lo terms, 3 also mid*mid): the LAST
      // t.split entries of the smallest-first list
      const int pass0 = t.split ? 6 - t.split : 5;
      // the whole warp walks the ring (warp-uniform state); one elected lane issues
      for (int pass = pass0; pass < 6; ++pass) {
        // split operands: this pass multiplies part pa of A with part pb of B (offsets along the concatenated dimension)
        const int oa = t.split ? tc_split_a(pass) * t.a_part : 0, ob = t.split ? tc_split_b(pass) * t.b_part : 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (elect_one()) {
            const uint32_t fb = CG == 2 ? mapa_shared(full_bar(stage), 0) : full_bar(stage);
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES * CG);
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            if constexpr (!MNMAJ) {
              tma_load_2d<CG>(sa, &t.tmA, fb, oa + kb * TC_BLOCK_K, m0);
              tma_load_2d<CG>(sa + TC_A_BYTES, &t.tmB, fb, ob + kb * TC_BLOCK_K, n0);
            } else {  // boxes of {64 M/N elements, 64 K rows}
#pragma unroll
              for (int i = 0; i < TC_BLOCK_M / 64; ++i)
                tma_load_2d<CG>(sa + i * TC_MN_BOX_BYTES, &t.tmA, fb, oa + m0 + 64 * i, kb * TC_BLOCK_K);
#pragma unroll
              for (int i = 0; i < Cfg::B_ROWS / 64; ++i)
                tma_load_2d<CG>(sa + TC_A_BYTES + i * TC_MN_BOX_BYTES, &t.tmB, fb, ob + n0 + 64 * i, kb * TC_BLOCK_K);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TC_BLOCK_M * CG, BN, MNMAJ);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = unit; tile < g.total_tiles; tile += nunits) {
        const int ti = tc_find_task(g, tile);
        const int nkb = ((g.t[ti].K + TC_BLOCK_K - 1) / TC_BLOCK_K) * (g.t[ti].split ? g.t[ti].split : 1);  // k-blocks over all passes
        mbar_wait(tempty_bar(as), aphase ^ 1u);  // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            const uint64_t da = MNMAJ ? umma_desc_mnmajor_sw128(sa) : umma_desc_kmajor_sw128(sa);
            const uint64_t db = MNMAJ ? umma_desc_mnmajor_sw128(sa + TC_A_BYTES) : umma_desc_kmajor_sw128(sa + TC_A_BYTES);
            // advancing K by 16: K-major: +32 B inside the 128-byte swizzle row; MN-major: +16 rows of 128 B
            constexpr uint64_t kstep = MNMAJ ? (uint64_t)((TC_UMMA_K * 128) >> 4) : 2u;
#pragma unroll
            for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k)
              umma_f16<CG>(d_tmem, da + kstep * (uint64_t)k, db + kstep * (uint64_t)k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit<CG>(empty_bar(stage));  // ring slot reusable once these MMAs have read it
            if (kb == nkb - 1) umma_commit<CG>(tfull_bar(as));  // accumulator complete
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++as == ACC) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int ew = warp - 2;          // 0..EW-1
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int part = ew >> 2;         // which column part of the 256-wide tile
    using Ep = TcEpi<EW, BN>;
    const int pc0 = Ep::part_begin(part) * 16, np = Ep::part_begin(part + 1) - Ep::part_begin(part);
    float* xp = xpose_ptr + ew * (TC_XPOSE_BYTES / 4);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = unit; tile < g.total_tiles; tile += nunits) {
      const int ti = tc_find_task(g, tile);
      const TcTask& t = g.t[ti];
      const Epi e = epi_pin(t.e);  // registers, not runtime-indexed parameter space (see epi_pin)
      const int local = tile - t.tile_begin;
      const int m0 = (local / t.tiles_n) * (TC_BLOCK_M * CG) + (int)rank * TC_BLOCK_M, n0 = (local % t.tiles_n) * BN;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + pc0);
      // (waits for the accumulator inside, after the first patch's loads are in flight)
      float r = tc_epi_warp<MODE, SUB, ACT, ELSRC, PLAIN, EW, BN>(e, m0 + q * 32, n0 + pc0, lane, t_row, xp, tfull_bar(as), aphase, np);
      // all of this warp's TMEM reads are complete (tcgen05.wait::ld above): release the stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(as), 0));
        else mbar_arrive(tempty_bar(as));
      }
      if (e.red != nullptr) {
        r = warp_sum(r);
        if (lane == 0) atomicAdd(e.red, r);
      }
      if (++as == ACC) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, TC_TMEM_COLS);
  }
}

}  // namespace jpcb
