// tc_staged.cuh -- the tcgen05 grouped GEMM of tc_gemm.cuh with a TMA-STAGED, ROW-OWNED fused epilogue, for the two
// phases that make up the inference loop of the wide bf16 configurations (BASELINE config 4):
//
//   ERR   e      = T - a D                     (prediction error; bf16 operand copy out, optional layer energy)
//   GRAD  z_new  = z - c (e_l - act'(z) (a D)) (activity gradient + Euler update; fp32 z_new and bf16 act(z_new) out)
//
// Why a second epilogue.  In tc_gemm.cuh the epilogue warps pull their inputs (z, e_l) with LDG into a register ring and
// push results with STG, after transposing the accumulator through shared memory so that those accesses coalesce.  ncu on
// the GRAD phase (profiles/r01_ncu_full_grad_final.txt): tensor pipe 28 % active, issue slots 19 %, the top stalls are the
// first use of a loaded z (long scoreboard) and the LDS of the transpose -- a latency-bound epilogue, 196 us against
// ~150 us of MMA time.  Here
//   * a LOADER warp streams the epilogue inputs of tile i (while its MMAs still run) into shared-memory SLOTS with TMA:
//     one slot = a 128-row x 32-column slab: fp32 z (or T) as 128-byte rows, SWIZZLE_128B, + bf16 e_l as 64-byte rows,
//     SWIZZLE_64B;  4 slots = up to 96 KB in flight per SM, no registers, no LSU instructions;
//   * the epilogue warps keep the TMEM ownership lane = row all the way: tcgen05.ld 32 columns of the row, LDS.128 of the
//     same row's chunks from the slot (the TMA swizzle makes a quarter-warp's 8 rows hit 8 different 16-byte bank groups:
//     conflict-free), arithmetic, STS.128 of the results IN PLACE (z_new over z, act(z_new) over e_l) -- no transpose;
//   * a STORER warp writes finished slabs back with TMA stores (cp.async.bulk.tensor ... global.shared::cta) and recycles
//     the slot once the store has read it.
// Hand-offs are mbarriers only (sfull: TMA bytes landed; sready: 128 row-owners done + proxy fence; sfree: store drained).
//
// Roles (384 threads, one CTA per SM, CTA pairs own 256 x 256 tiles exactly as in tc_gemm.cuh):
//   warp 0  mainloop TMA producer      warp 1  MMA issuer (leader CTA)      warp 2  epilogue loader      warp 3  storer
//   warps 4..11  epilogue: two teams of 4 warps (one warp per TMEM lane quarter); slabs alternate between the teams.
#pragma once
#include "tc_gemm.cuh"

namespace jpcb {

enum : int { TS_ERR = 0, TS_GRAD = 1 };

constexpr int TS_STAGES = 4;                        // default mainloop ring: 4 x 32 KB
constexpr int TS_SLAB_COLS = 32;
constexpr int TS_SLAB_F32 = TC_BLOCK_M * TS_SLAB_COLS * 4;   // 16 KB
constexpr int TS_SLAB_B16 = TC_BLOCK_M * TS_SLAB_COLS * 2;   //  8 KB
constexpr int TS_SLOT_BYTES = TS_SLAB_F32 + TS_SLAB_B16;     // 24 KB
constexpr int TS_SLOTS = 4;
constexpr int TS_EW = 8;
constexpr int TS_THREADS = 32 * (4 + TS_EW);
constexpr int TS_TEAM_THREADS = 128;
constexpr int TS_BAR_BYTES = 256;
constexpr int ts_smem(int stages, int slots) { return 1024 + stages * TcCfg<2, 256>::STAGE_BYTES + slots * TS_SLOT_BYTES + TS_BAR_BYTES; }
constexpr int TS_SMEM = ts_smem(TS_STAGES, TS_SLOTS);
static_assert(TS_SMEM <= 227 * 1024, "staged kernel: shared memory over the per-CTA limit");

struct alignas(64) TsTask {
  CUtensorMap tmA;   // bf16 [M, K] row-major, box {64, 128}
  CUtensorMap tmB;   // bf16 [N, K] row-major, box {64, 128}
  CUtensorMap tmI0;  // fp32 [M, N], box {32, 128}, SWIZZLE_128B: GRAD z / ERR target T
  CUtensorMap tmI1;  // bf16 [M, N], box {32, 128}, SWIZZLE_64B : GRAD e_l
  CUtensorMap tmO0;  // fp32 [M, N]: GRAD z_new
  CUtensorMap tmO1;  // bf16 [M, N]: GRAD act(z_new) / ERR e
  int K, tile_begin, tiles_n, kind;
  int M, N;
  float alpha;  // layer scaling a_l
  float cg;     // GRAD: dt / B
  float* red;   // ERR: layer-energy accumulator or null
  int pad_[6];
};

constexpr int TS_MAX_TASKS = 16;
struct alignas(64) TsGroup {
  int ntasks;
  int total_tiles;
  int pf_dist;  // > 0: the mainloop producer prefetches its A blocks into L2 this many k-blocks ahead of the ring
  int hints;    // bit 0: evict_first on the fp32 slab loads, bit 1: evict_first on the fp32 slab stores, bit 2: evict_last on B (weights)
  int dbg;      // measurement only (results are garbage): bit 0 no slab loads, 1 no slab stores, 2 no epilogue arithmetic, 3 no mainloop loads
  int pad_[11];
  TsTask t[TS_MAX_TASKS];
};

__device__ __forceinline__ int ts_find_task(const TsGroup& g, int tile) {
  int ti = 0;
  while (ti + 1 < g.ntasks && tile >= g.t[ti + 1].tile_begin) ++ti;
  return ti;
}

// ---- PTX: TMA stores, proxy fence, vector shared-memory access ----
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, uint32_t src, int c_inner, int c_outer, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c_inner), "r"(c_outer), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c_inner, int c_outer, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c_inner), "r"(c_outer), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c_inner, int c_outer, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c_inner), "r"(c_outer), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory (STS) -> visible to the async proxy (the TMA store that reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// act'(z) and act(z) of the two activations the staged variants are compiled for (fast forms: the operands of this path
// are bf16 anyway, same as the specialised variants of tc_gemm.cuh)
template <int ACT>
__device__ __forceinline__ float ts_act(float x) {
  if constexpr (ACT == ACT_TANH) return tanh_f<true>(x);
  else if constexpr (ACT == ACT_RELU) return x > 0.f ? x : 0.f;
  else return x;
}
template <int ACT>
__device__ __forceinline__ float ts_dact(float x) {
  if constexpr (ACT == ACT_TANH) { const float t = tanh_f<true>(x); return fmaf(-t, t, 1.f); }
  else if constexpr (ACT == ACT_RELU) return x > 0.f ? 1.f : 0.f;
  else return 1.f;
}

// One 128-row x 32-column slab, this thread's row.  `sz`: the row's 128 bytes in the fp32 region of the slot, `se`: its 64
// bytes in the bf16 region; `xz` / `xe`: the row's XOR terms of the two TMA swizzles (row & 7, (row >> 1) & 3).
template <int ACT>
__device__ __forceinline__ void ts_slab_grad(const uint32_t (&acc)[32], uint32_t sz, uint32_t se, uint32_t xz, uint32_t xe, float alpha,
                                             float cg) {
#pragma unroll
  for (int h = 0; h < 4; ++h) {  // 8 columns: two fp32 chunks, one bf16 chunk
    const uint32_t a0 = sz + (((uint32_t)(2 * h) ^ xz) << 4), a1 = sz + (((uint32_t)(2 * h + 1) ^ xz) << 4);
    const uint32_t ae = se + (((uint32_t)h ^ xe) << 4);
    const float4 z0 = lds128(a0), z1 = lds128(a1);
    const float4 ev = lds128(ae);
    const float z[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
    const uint32_t eu[4] = {__float_as_uint(ev.x), __float_as_uint(ev.y), __float_as_uint(ev.z), __float_as_uint(ev.w)};
    float zn[8], ph[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float el = (k & 1) ? bf16_hi(eu[k >> 1]) : bf16_lo(eu[k >> 1]);
      const float g = fmaf(-ts_dact<ACT>(z[k]), alpha * __uint_as_float(acc[8 * h + k]), el);
      zn[k] = fmaf(-cg, g, z[k]);
      ph[k] = ts_act<ACT>(zn[k]);
    }
    sts128(a0, zn[0], zn[1], zn[2], zn[3]);
    sts128(a1, zn[4], zn[5], zn[6], zn[7]);
    sts128u(ae, pack_bf16(ph[0], ph[1]), pack_bf16(ph[2], ph[3]), pack_bf16(ph[4], ph[5]), pack_bf16(ph[6], ph[7]));
  }
}

// ERR slab: e = T - a D -> bf16; returns 1/2 sum e^2 over the row's VALID columns (`ncols` of the slab are inside the matrix)
__device__ __forceinline__ float ts_slab_err(const uint32_t (&acc)[32], uint32_t sz, uint32_t se, uint32_t xz, uint32_t xe, float alpha,
                                             bool want_red, int ncols) {
  float r = 0.f;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float4 t0 = lds128(sz + (((uint32_t)(2 * h) ^ xz) << 4)), t1 = lds128(sz + (((uint32_t)(2 * h + 1) ^ xz) << 4));
    const float t[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    float e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = fmaf(-alpha, __uint_as_float(acc[8 * h + k]), t[k]);
    if (want_red) {
#pragma unroll
      for (int k = 0; k < 8; ++k) r += (8 * h + k < ncols) ? 0.5f * e[k] * e[k] : 0.f;
    }
    sts128u(se + (((uint32_t)h ^ xe) << 4), pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
  }
  return r;
}

template <int ACT, int STAGES = TS_STAGES, int SLOTS = TS_SLOTS>
__global__ void __launch_bounds__(TS_THREADS, 1) tc_staged_gemm(const __grid_constant__ TsGroup g) {
  constexpr int CG = 2, BN = TC_BLOCK_N;
  using Cfg = TcCfg<CG, BN>;
  constexpr int ACC = TC_TMEM_COLS / BN;
  constexpr int TS_SLOTS = SLOTS;  // (shadows the default slot count)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t slot_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = slot_base + TS_SLOTS * TS_SLOT_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC + s); };
  auto sfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 * ACC + s); };
  auto sready_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 * ACC + TS_SLOTS + s); };
  auto sfree_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 * ACC + 2 * TS_SLOTS + s); };
  constexpr int NBAR = 2 * STAGES + 2 * ACC + 3 * TS_SLOTS;
  static_assert(8 * NBAR + 4 <= TS_BAR_BYTES, "barrier block exceeds its slack");
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bar_base - smem_base) + 8 * NBAR);
  const uint32_t tmem_slot_addr = bar_base + 8u * NBAR;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int unit = (int)blockIdx.x / CG, nunits = (int)gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < ACC; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), TS_EW * CG); }
    for (int s = 0; s < TS_SLOTS; ++s) { mbar_init(sfull_bar(s), 1); mbar_init(sready_bar(s), TS_TEAM_THREADS); mbar_init(sfree_bar(s), 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<CG>(tmem_slot_addr, TC_TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===================== mainloop TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    // L2 prefetch stream of this CTA's A blocks, pf_dist k-blocks ahead of the ring (across tile boundaries)
    const int pf_dist = g.pf_dist;
    int pf_tile = unit, pf_kb = 0, pf_nkb = 0, pf_m0 = 0;
    const CUtensorMap* pf_map = nullptr;
    auto pf_open = [&]() {
      if (pf_tile < g.total_tiles) {
        const TsTask& t = g.t[ts_find_task(g, pf_tile)];
        pf_map = &t.tmA;
        pf_nkb = (t.K + TC_BLOCK_K - 1) / TC_BLOCK_K;
        pf_m0 = ((pf_tile - t.tile_begin) / t.tiles_n) * (TC_BLOCK_M * CG) + (int)rank * TC_BLOCK_M;
        pf_kb = 0;
      }
    };
    auto pf_step = [&]() {  // (warp-uniform state; one lane issues)
      if (pf_tile >= g.total_tiles) return;
      if (elect_one()) tma_prefetch_2d(pf_map, pf_kb * TC_BLOCK_K, pf_m0);
      __syncwarp();
      if (++pf_kb == pf_nkb) { pf_tile += nunits; pf_open(); }
    };
    if (pf_dist > 0) {
      pf_open();
      for (int i = 0; i < pf_dist; ++i) pf_step();
    }
    const uint64_t pol_b = policy_evict_last();
    const bool hint_b = (g.hints & 4) != 0;
    for (int tile = unit; tile < g.total_tiles; tile += nunits) {
      const TsTask& t = g.t[ts_find_task(g, tile)];
      const int local = tile - t.tile_begin;
      const int m0 = (local / t.tiles_n) * (TC_BLOCK_M * CG) + (int)rank * TC_BLOCK_M;
      const int n0 = (local % t.tiles_n) * BN + (int)rank * Cfg::B_ROWS;
      const int nkb = (t.K + TC_BLOCK_K - 1) / TC_BLOCK_K;
      for (int kb = 0; kb < nkb; ++kb) {
        if (pf_dist > 0) pf_step();
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t fb = mapa_shared(full_bar(stage), 0);
          if (g.dbg & 8) { if (rank == 0) mbar_arrive(full_bar(stage)); }
          else {
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES * CG);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          tma_load_2d<CG>(sa, &t.tmA, fb, kb * TC_BLOCK_K, m0);
          if (hint_b) tma_load_2d_cg2_hint(sa + TC_A_BYTES, &t.tmB, fb, kb * TC_BLOCK_K, n0, pol_b);
          else tma_load_2d<CG>(sa + TC_A_BYTES, &t.tmB, fb, kb * TC_BLOCK_K, n0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TC_BLOCK_M * CG, BN, false);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = unit; tile < g.total_tiles; tile += nunits) {
        const int nkb = (g.t[ts_find_task(g, tile)].K + TC_BLOCK_K - 1) / TC_BLOCK_K;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            const uint64_t da = umma_desc_kmajor_sw128(sa), db = umma_desc_kmajor_sw128(sa + TC_A_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k)
              umma_f16<CG>(d_tmem, da + 2u * (uint64_t)k, db + 2u * (uint64_t)k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit<CG>(empty_bar(stage));
            if (kb == nkb - 1) umma_commit<CG>(tfull_bar(as));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++as == ACC) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp == 2) {
    // ===================== epilogue loader: slab inputs of this CTA's 128 rows, in tile order =====================
    uint32_t j = 0;  // running slab counter: slot = j % SLOTS, use parity = (j / SLOTS) & 1
    const uint64_t pol_f = policy_evict_first();
    const bool hint_ld = (g.hints & 1) != 0;
    for (int tile = unit; tile < g.total_tiles; tile += nunits) {
      const TsTask& t = g.t[ts_find_task(g, tile)];
      const int local = tile - t.tile_begin;
      const int m0_ = (local / t.tiles_n) * (TC_BLOCK_M * CG) + (int)rank * TC_BLOCK_M, n0 = (local % t.tiles_n) * BN;
      const int nsl = min(BN / TS_SLAB_COLS, (t.N - n0 + TS_SLAB_COLS - 1) / TS_SLAB_COLS);
      const bool grad = t.kind == TS_GRAD;
      for (int s = 0; s < nsl; ++s, ++j) {
        const int slot = (int)(j % TS_SLOTS);
        mbar_wait(sfree_bar(slot), ((j / TS_SLOTS) & 1u) ^ 1u);
        if (elect_one()) {
          const uint32_t dst = slot_base + slot * TS_SLOT_BYTES;
          const int m0 = (g.dbg & 32) ? (m0_ & 255) : m0_;
          if (g.dbg & 1) mbar_arrive(sfull_bar(slot));
          else {
          mbar_arrive_expect_tx(sfull_bar(slot), grad ? TS_SLOT_BYTES : TS_SLAB_F32);
          if (hint_ld) tma_load_2d_hint(dst, &t.tmI0, sfull_bar(slot), n0 + s * TS_SLAB_COLS, m0, pol_f);
          else tma_load_2d<1>(dst, &t.tmI0, sfull_bar(slot), n0 + s * TS_SLAB_COLS, m0);
          if (grad) tma_load_2d<1>(dst + TS_SLAB_F32, &t.tmI1, sfull_bar(slot), n0 + s * TS_SLAB_COLS, m0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ===================== storer: finished slabs back to HBM, slots recycled once the store has read them ==========
    uint32_t j = 0;
    int prev_slot = -1;
    const uint64_t pol_f = policy_evict_first();
    const bool hint_st = (g.hints & 2) != 0;
    for (int tile = unit; tile < g.total_tiles; tile += nunits) {
      const TsTask& t = g.t[ts_find_task(g, tile)];
      const int local = tile - t.tile_begin;
      const int m0s = (local / t.tiles_n) * (TC_BLOCK_M * CG) + (int)rank * TC_BLOCK_M, n0 = (local % t.tiles_n) * BN;
      const int nsl = min(BN / TS_SLAB_COLS, (t.N - n0 + TS_SLAB_COLS - 1) / TS_SLAB_COLS);
      const bool grad = t.kind == TS_GRAD;
      for (int s = 0; s < nsl; ++s, ++j) {
        const int slot = (int)(j % TS_SLOTS);
        mbar_wait(sready_bar(slot), (j / TS_SLOTS) & 1u);
        if (lane == 0) {  // (one fixed lane: bulk async-groups are per thread)
          const uint32_t src = slot_base + slot * TS_SLOT_BYTES;
          const int m0 = (g.dbg & 16) ? (m0s & 255) : m0s;
          if (g.dbg & 2) {
          } else if (grad) {
            if (hint_st) tma_store_2d_hint(&t.tmO0, src, n0 + s * TS_SLAB_COLS, m0, pol_f);
            else tma_store_2d(&t.tmO0, src, n0 + s * TS_SLAB_COLS, m0);
          }
          if (!(g.dbg & 2)) tma_store_2d(&t.tmO1, src + TS_SLAB_F32, n0 + s * TS_SLAB_COLS, m0);
          bulk_commit();
          if (prev_slot >= 0) {
            bulk_wait_read<1>();  // everything but the group just committed has been read out of shared memory
            mbar_arrive(sfree_bar(prev_slot));
          }
        }
        prev_slot = slot;
        __syncwarp();
      }
    }
    if (lane == 0) bulk_wait_all();  // all stores complete (written) before the CTA may exit
  } else {
    // ===================== epilogue warps: this CTA's 128 rows, lane = row =====================
    const int ew = warp - 4, team = ew >> 2, q = warp & 3;
    const uint32_t row = (uint32_t)(q * 32 + lane);
    const uint32_t xz = row & 7u, xe = (row >> 1) & 3u;
    int as = 0;
    uint32_t aphase = 0, j = 0;
    for (int tile = unit; tile < g.total_tiles; tile += nunits) {
      const TsTask& t = g.t[ts_find_task(g, tile)];
      const int local = tile - t.tile_begin;
      const int n0 = (local % t.tiles_n) * BN;
      const int nsl = min(BN / TS_SLAB_COLS, (t.N - n0 + TS_SLAB_COLS - 1) / TS_SLAB_COLS);
      const int kind = t.kind;
      const float alpha = t.alpha, cg = t.cg;
      float* const red = t.red;
      const int m0 = (local / t.tiles_n) * (TC_BLOCK_M * CG) + (int)rank * TC_BLOCK_M;
      const bool row_ok = m0 + (int)row < t.M;
      const int ncol_tile = t.N - n0;
      float r = 0.f;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      for (int s = 0; s < nsl; ++s, ++j) {
        if ((int)(j & 1u) != team) continue;
        const int slot = (int)(j % TS_SLOTS);
        const uint32_t sb = slot_base + slot * TS_SLOT_BYTES;
        const uint32_t sz = sb + row * (TS_SLAB_COLS * 4), se = sb + TS_SLAB_F32 + row * (TS_SLAB_COLS * 2);
        uint32_t acc[32];
        tmem_ld32(t_row + (uint32_t)(s * TS_SLAB_COLS), acc);
        mbar_wait(sfull_bar(slot), (j / TS_SLOTS) & 1u);
        tmem_ld_wait();
        if (g.dbg & 4) {
          if (acc[0] == 0x7fc12345u) sts128u(sz, acc[1], acc[2], acc[3], acc[4]);  // (keeps the TMEM load alive)
        } else if (kind == TS_GRAD) ts_slab_grad<ACT>(acc, sz, se, xz, xe, alpha, cg);
        else {
          const float rr = ts_slab_err(acc, sz, se, xz, xe, alpha, red != nullptr, ncol_tile - s * TS_SLAB_COLS);
          r += row_ok ? rr : 0.f;
        }
        fence_async_smem();
        mbar_arrive(sready_bar(slot));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(as), 0));
      if (red != nullptr) {
        r = warp_sum(r);
        if (lane == 0) atomicAdd(red, r);
      }
      if (++as == ACC) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, TC_TMEM_COLS);
  }
}

}  // namespace jpcb
