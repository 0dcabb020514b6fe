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
// Whole-solve schedule ("persist").  One launch can run ALL T Euler steps of the inference loop: the tile stream is
//   step 0: ERR tiles, GRAD tiles; step 1: ERR tiles, GRAD tiles; ...   (within a phase: layer, row block, column tile)
// dealt round-robin to the CTA pairs, which walk it in order.  A tile reads what tiles of the previous phase wrote for the
// same 256-row block (samples are independent): before its first global read the producer / loader warps wait until
// per-(task, row block) completion counters in global memory reach the value the schedule implies; the storer warp bumps
// the tile's counter once its TMA stores have completed.  Because a phase is ~12 rounds of the grid long, the counters a
// tile waits for were normally reached long before -- the waits are there for correctness, the gain is that 2T launches
// with their wave tails, the separate layer-1 launch and the per-launch tile quantisation (13 + 11 + 2 rounds for 24.2
// rounds of work on config 4) collapse into one continuous stream.
//
// Roles (384 threads, one CTA per SM, CTA pairs own 256 x 256 tiles exactly as in tc_gemm.cuh):
//   warp 0  mainloop TMA producer      warp 1  MMA issuer (leader CTA)      warp 2  epilogue loader      warp 3  storer
//   warps 4..11  epilogue: two teams of 4 warps (one warp per TMEM lane quarter); slabs alternate between the teams.
#pragma once
#include "tc_gemm.cuh"

namespace jpcb {

enum : int { TS_ERR = 0, TS_GRAD = 1, TS_GRAD_P = 2 };

constexpr int TS_STAGES = 4;                        // mainloop ring: 4 x 32 KB
constexpr int TS_SLAB_COLS = 32;                    // ERR / GRAD slabs; GRAD_P (e_l = z - P, two fp32 inputs) uses 16
constexpr int TS_SLAB_F32 = TC_BLOCK_M * TS_SLAB_COLS * 4;   // 16 KB
constexpr int TS_SLAB_B16 = TC_BLOCK_M * TS_SLAB_COLS * 2;   //  8 KB
constexpr int TS_SLOT_BYTES = TS_SLAB_F32 + TS_SLAB_B16;     // 24 KB: [fp32 region 16 KB | bf16 region 8 KB]
constexpr int TS_SLOTS = 4;
constexpr int TS_EW = 8;
constexpr int TS_THREADS = 32 * (4 + TS_EW);
constexpr int TS_TEAM_THREADS = 128;
constexpr int TS_BAR_BYTES = 256;
constexpr int ts_smem(int stages, int slots) { return 1024 + stages * TcCfg<2, 256>::STAGE_BYTES + slots * TS_SLOT_BYTES + TS_BAR_BYTES; }
constexpr int TS_SMEM = ts_smem(TS_STAGES, TS_SLOTS);
static_assert(TS_SMEM <= 227 * 1024, "staged kernel: shared memory over the per-CTA limit");

// one dependency of a task in the whole-solve schedule: counter row `cnt[r]` must have reached
// max(0, step - a + off) * n * 2 (n column tiles per row block and step, two CTAs per tile; `a`: first step the producing
// task is active in; off = 1: produced in the same step (ERR -> GRAD), 0: in the previous one (GRAD -> next ERR))
struct TsDep {
  const int* cnt;
  int n, a, off, pad_;
};

struct alignas(64) TsTask {
  CUtensorMap tmA;   // bf16 [M, K] row-major, box {64, 128}
  CUtensorMap tmB;   // bf16 [N, K] row-major, box {64, 128}
  CUtensorMap tmI0;  // fp32 [M, N], box {32 | 16, 128}: GRAD z / ERR target T
  CUtensorMap tmI1;  // GRAD: bf16 e_l [M, N], box {32, 128};  GRAD_P: fp32 hoisted prediction P, box {16, 128}
  CUtensorMap tmO0;  // fp32 [M, N]: GRAD z_new
  CUtensorMap tmO1;  // bf16 [M, N]: GRAD act(z_new) / ERR e
  int K, tile_begin, tiles_n, kind;
  int M, N;
  float alpha;    // layer scaling a_l
  float cg;       // GRAD: dt / B
  float cg_last;  // GRAD: the same for the last (clipped) step of a whole-solve schedule
  int pad0_;
  float* red;     // ERR: layer-energy accumulator or null
  int* done;      // whole-solve schedule: completion counters of this task, one per row block
  float* o0;            // the tensors behind tmO0 / tmO1 (the storer's LSU path, TsGroup::stg)
  __nv_bfloat16* o1;
  const float* bias;    // ERR: per-column bias of the layer (e = T - a (D + b)) or null
  TsDep dep[2];
};

constexpr int TS_MAX_TASKS = 32;
constexpr int TS_MAX_WARM = 16;
struct alignas(64) TsGroup {
  int ntasks;
  int total_tiles;
  // ---- whole-solve schedule (persist != 0): tasks [0, L1) are the ERR tasks of layers 1..L1, [L1, 2 L1) the GRAD tasks ----
  int persist;
  int R;           // row blocks of 256 rows
  int nsteps;
  int L1;
  int wave;        // error wavefront: at step s only layer tasks i >= max(0, L1 - 1 - s) are active
  int warm;        // steps with fewer active layers than L1 (0 without wavefront)
  int tiles_full;  // tiles of a step with every layer active
  int stg;         // storer writes finished slabs with coalesced st.global (LSU) instead of TMA stores
  int warm_begin[TS_MAX_WARM + 1];  // first tile of warm-up step s; warm_begin[warm] = first tile of the first full step
  int hints;       // L2 eviction-priority hints: bit 0 weights (B operand) evict_last, bit 1 slab loads evict_first, bit 2 slab stores evict_first
  int pad_[4];
  TsTask t[TS_MAX_TASKS];
};

__device__ __forceinline__ int ts_find_task(const TsGroup& g, int tile) {
  int ti = 0;
  while (ti + 1 < g.ntasks && tile >= g.t[ti + 1].tile_begin) ++ti;
  return ti;
}

// tile of the launch -> task, row block, first column, step
struct TsTile {
  int ti, r, n0, step;
};
__device__ __forceinline__ TsTile ts_decode(const TsGroup& g, int tile) {
  TsTile o;
  if (!g.persist) {
    o.ti = ts_find_task(g, tile);
    const TsTask& t = g.t[o.ti];
    const int local = tile - t.tile_begin;
    o.r = local / t.tiles_n;
    o.n0 = (local % t.tiles_n) * TC_BLOCK_N;
    o.step = 0;
    return o;
  }
  int s = 0, off;
  if (tile < g.warm_begin[g.warm]) {
    while (tile >= g.warm_begin[s + 1]) ++s;
    off = tile - g.warm_begin[s];
  } else {
    const int q = tile - g.warm_begin[g.warm];
    s = g.warm + q / g.tiles_full;
    off = q % g.tiles_full;
  }
  const int L1 = g.L1;
  const int i0 = g.wave ? max(0, L1 - 1 - s) : 0;
  // inside a phase: layer, row block, column tile (the tiles in flight at any time share one or two layers' weights in L2)
  int base = 0, i = i0;
  for (;;) {
    const int nt = g.R * g.t[base + i].tiles_n;
    if (off < nt) break;
    off -= nt;
    if (++i == L1) { i = i0; base = L1; }  // ERR phase done: GRAD phase of the step
  }
  o.ti = base + i;
  o.r = off / g.t[o.ti].tiles_n;
  o.n0 = (off - o.r * g.t[o.ti].tiles_n) * TC_BLOCK_N;
  o.step = s;
  return o;
}
__device__ __forceinline__ int ts_slab_cols(int kind) { return kind == TS_GRAD_P ? 16 : TS_SLAB_COLS; }
__device__ __forceinline__ int ts_num_slabs(const TsTask& t, int n0) {
  const int sc = ts_slab_cols(t.kind);
  return (min(TC_BLOCK_N, t.N - n0) + sc - 1) / sc;
}

// ---- PTX: TMA stores, proxy fences, vector shared-memory access, release / acquire counters ----
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c_inner), "r"(c_outer)
               : "memory");
}
// L2 eviction-priority policies and the hinted forms of the TMA instructions
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_cg2_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c_inner, int c_outer, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c_inner), "r"(c_outer), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c_inner, int c_outer, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c_inner), "r"(c_outer), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, uint32_t src, int c_inner, int c_outer, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c_inner), "r"(c_outer), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory (STS) -> visible to the async proxy (the TMA store that reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// orders this thread's generic-proxy and async-proxy (TMA) accesses to global memory
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
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

// Whole-solve schedule: every lane of the calling warp waits (acquire, gpu scope) until the tiles this one reads from are
// complete, then orders its coming TMA (async-proxy) reads after what it acquired.  Bounded like mbar_wait.
__device__ __forceinline__ void ts_wait_deps(const TsTask& t, int step, int r) {
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const TsDep& dp = t.dep[d];
    if (dp.cnt == nullptr) continue;
    const int target = max(0, step - dp.a + dp.off) * dp.n * 2;
    if (target <= 0) continue;
    const int* p = dp.cnt + r;
    if (ld_acquire_gpu(p) >= target) continue;
    uint64_t t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint32_t spins = 0;
    while (ld_acquire_gpu(p) < target) {
      __nanosleep(64);
      if ((++spins & 1023u) == 0) {
        uint64_t t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) __trap();
      }
    }
  }
  fence_async_global();
}

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

// One 128-row x 32-column slab, this thread's row.  `sb`: the slot; fp32 region: 128-byte rows, SWIZZLE_128B (16-byte
// chunk c of row r sits at chunk c ^ (r & 7)); bf16 region at +16 KB: 64-byte rows, SWIZZLE_64B (c ^ ((r >> 1) & 3)).
template <int ACT>
__device__ __forceinline__ void ts_slab_grad(const uint32_t (&acc)[32], uint32_t sb, uint32_t row, float alpha, float cg) {
  const uint32_t sz = sb + row * 128u, se = sb + TS_SLAB_F32 + row * 64u, xz = row & 7u, xe = (row >> 1) & 3u;
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

// GRAD_P slab (16 columns): e_l = z - P with the hoisted layer-0 prediction P (fp32).  Slot: z at +0 and P at +8 KB as
// 64-byte rows, SWIZZLE_64B; outputs z_new in place and act(z_new) at +16 KB as 32-byte rows, SWIZZLE_32B (c ^ ((r >> 2) & 1)).
template <int ACT>
__device__ __forceinline__ void ts_slab_grad_p(const uint32_t (&acc)[16], uint32_t sb, uint32_t row, float alpha, float cg) {
  const uint32_t sz = sb + row * 64u, sp = sb + 8192u + row * 64u, so = sb + TS_SLAB_F32 + row * 32u;
  const uint32_t x4 = (row >> 1) & 3u, x2 = (row >> 2) & 1u;
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // 8 columns
    const uint32_t a0 = sz + (((uint32_t)(2 * h) ^ x4) << 4), a1 = sz + (((uint32_t)(2 * h + 1) ^ x4) << 4);
    const float4 z0 = lds128(a0), z1 = lds128(a1);
    const float4 p0 = lds128(sp + (((uint32_t)(2 * h) ^ x4) << 4)), p1 = lds128(sp + (((uint32_t)(2 * h + 1) ^ x4) << 4));
    const float z[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
    const float pr[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float zn[8], ph[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float g = fmaf(-ts_dact<ACT>(z[k]), alpha * __uint_as_float(acc[8 * h + k]), z[k] - pr[k]);
      zn[k] = fmaf(-cg, g, z[k]);
      ph[k] = ts_act<ACT>(zn[k]);
    }
    sts128(a0, zn[0], zn[1], zn[2], zn[3]);
    sts128(a1, zn[4], zn[5], zn[6], zn[7]);
    sts128u(so + (((uint32_t)h ^ x2) << 4), pack_bf16(ph[0], ph[1]), pack_bf16(ph[2], ph[3]), pack_bf16(ph[4], ph[5]), pack_bf16(ph[6], ph[7]));
  }
}

// ERR slab: e = T - a D -> bf16; returns 1/2 sum e^2 over the row's VALID columns (`ncols` of the slab are inside the matrix)
// `bias`: the layer's bias at the slab's first column, or null (widths are multiples of 8: a group of 8 columns is inside or
// outside the matrix as a whole)
__device__ __forceinline__ float ts_slab_err(const uint32_t (&acc)[32], uint32_t sb, uint32_t row, float alpha, bool want_red, int ncols,
                                             const float* __restrict__ bias) {
  const uint32_t sz = sb + row * 128u, se = sb + TS_SLAB_F32 + row * 64u, xz = row & 7u, xe = (row >> 1) & 3u;
  float r = 0.f;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float4 t0 = lds128(sz + (((uint32_t)(2 * h) ^ xz) << 4)), t1 = lds128(sz + (((uint32_t)(2 * h + 1) ^ xz) << 4));
    const float t[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (bias != nullptr && 8 * h < ncols) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + 8 * h)), b1 = __ldg(reinterpret_cast<const float4*>(bias + 8 * h + 4));
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
    }
    float e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = fmaf(-alpha, __uint_as_float(acc[8 * h + k]) + b[k], t[k]);
    if (want_red) {
#pragma unroll
      for (int k = 0; k < 8; ++k) r += (8 * h + k < ncols) ? 0.5f * e[k] * e[k] : 0.f;
    }
    sts128u(se + (((uint32_t)h ^ xe) << 4), pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
  }
  return r;
}

// Storer, LSU path: one warp copies a finished slab (this CTA's 128 rows) from its slot to global memory with 16-byte
// accesses -- a row of the slab is `CH` 16-byte chunks, 32 / CH rows per instruction, chunk c of row r sits at c ^ swz(r).
template <int CH>
__device__ __forceinline__ void ts_stg_region(uint32_t src, uint8_t* gbase, size_t row_pitch_bytes, int rows_valid, int bytes_valid,
                                              int lane) {
  constexpr int RPI = 32 / CH;  // rows per instruction
  const int c = lane % CH, rl = lane / CH;
#pragma unroll 4
  for (int it = 0; it < TC_BLOCK_M / RPI; ++it) {
    const int r = it * RPI + rl;
    const uint32_t sw = CH == 8 ? (uint32_t)(r & 7) : (CH == 4 ? (uint32_t)((r >> 1) & 3) : (uint32_t)((r >> 2) & 1));
    const float4 v = lds128(src + (uint32_t)r * (16u * CH) + (((uint32_t)c ^ sw) << 4));
    if (r < rows_valid && 16 * c < bytes_valid)   // (widths are multiples of 8 elements: a chunk is inside or outside as a whole)
      __stcg(reinterpret_cast<float4*>(gbase + (size_t)r * row_pitch_bytes + 16 * c), v);
  }
}

template <int ACT, int STAGES = TS_STAGES, int SLOTS = TS_SLOTS>
__global__ void __launch_bounds__(TS_THREADS, 1) tc_staged_gemm(const __grid_constant__ TsGroup g) {
  constexpr int CG = 2, BN = TC_BLOCK_N;
  using Cfg = TcCfg<CG, BN>;
  constexpr int ACC = TC_TMEM_COLS / BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t slot_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = slot_base + SLOTS * TS_SLOT_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC + s); };
  auto sfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 * ACC + s); };
  auto sready_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 * ACC + SLOTS + s); };
  auto sfree_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 * ACC + 2 * SLOTS + s); };
  constexpr int NBAR = 2 * STAGES + 2 * ACC + 3 * SLOTS;
  static_assert(8 * NBAR + 4 <= TS_BAR_BYTES, "barrier block exceeds its slack");
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bar_base - smem_base) + 8 * NBAR);
  const uint32_t tmem_slot_addr = bar_base + 8u * NBAR;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int unit = (int)blockIdx.x / CG, nunits = (int)gridDim.x / CG;
  const bool persist = g.persist != 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < ACC; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), TS_EW * CG); }
    for (int s = 0; s < SLOTS; ++s) { mbar_init(sfull_bar(s), 1); mbar_init(sready_bar(s), TS_TEAM_THREADS); mbar_init(sfree_bar(s), 1); }
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
    const uint64_t pol_last = l2_policy_evict_last();
    for (int tile = unit; tile < g.total_tiles; tile += nunits) {
      const TsTile ti = ts_decode(g, tile);
      const TsTask& t = g.t[ti.ti];
      const int m0 = ti.r * (TC_BLOCK_M * CG) + (int)rank * TC_BLOCK_M;
      const int n0 = ti.n0 + (int)rank * Cfg::B_ROWS;
      const int nkb = (t.K + TC_BLOCK_K - 1) / TC_BLOCK_K;
      if (persist) ts_wait_deps(t, ti.step, ti.r);  // (the A operand of this tile is another tile's output)
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t fb = mapa_shared(full_bar(stage), 0);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES * CG);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          tma_load_2d<CG>(sa, &t.tmA, fb, kb * TC_BLOCK_K, m0);
          if (g.hints & 1) tma_load_2d_cg2_hint(sa + TC_A_BYTES, &t.tmB, fb, kb * TC_BLOCK_K, n0, pol_last);
          else tma_load_2d<CG>(sa + TC_A_BYTES, &t.tmB, fb, kb * TC_BLOCK_K, n0);
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
        const int nkb = (g.t[ts_decode(g, tile).ti].K + TC_BLOCK_K - 1) / TC_BLOCK_K;
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
    const uint64_t pol_first = l2_policy_evict_first();
    for (int tile = unit; tile < g.total_tiles; tile += nunits) {
      const TsTile ti = ts_decode(g, tile);
      const TsTask& t = g.t[ti.ti];
      const int m0 = ti.r * (TC_BLOCK_M * CG) + (int)rank * TC_BLOCK_M, n0 = ti.n0;
      const int nsl = ts_num_slabs(t, n0), sc = ts_slab_cols(t.kind), kind = t.kind;
      if (persist) ts_wait_deps(t, ti.step, ti.r);
      for (int s = 0; s < nsl; ++s, ++j) {
        const int slot = (int)(j % SLOTS);
        mbar_wait(sfree_bar(slot), ((j / SLOTS) & 1u) ^ 1u);
        if (elect_one()) {
          const uint32_t dst = slot_base + slot * TS_SLOT_BYTES, bar = sfull_bar(slot);
          const bool hf = (g.hints & 2) != 0;
          auto ld = [&](uint32_t d_, const CUtensorMap* mp) {
            if (hf) tma_load_2d_hint(d_, mp, bar, n0 + s * sc, m0, pol_first);
            else tma_load_2d<1>(d_, mp, bar, n0 + s * sc, m0);
          };
          if (kind == TS_GRAD_P) {
            mbar_arrive_expect_tx(bar, 16384);
            ld(dst, &t.tmI0);
            ld(dst + 8192, &t.tmI1);
          } else {
            mbar_arrive_expect_tx(bar, kind == TS_GRAD ? TS_SLOT_BYTES : TS_SLAB_F32);
            ld(dst, &t.tmI0);
            if (kind == TS_GRAD) ld(dst + TS_SLAB_F32, &t.tmI1);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ===================== storer: finished slabs back to HBM, slots recycled once the store has read them ==========
    uint32_t j = 0;
    int prev_slot = -1;
    // whole-solve schedule: a finished tile is published (its row block's counter bumped) once its stores have COMPLETED;
    // rather than stalling for that, the check rides on the stores of the next tile: after TS_PUB_LAG more groups have
    // been committed, "at most TS_PUB_LAG groups pending" implies every group of the previous tile is complete
    constexpr int TS_PUB_LAG = 4;
    int* pend = nullptr;
    int since = 0;
    auto publish = [&]() {  // (lane 0)
      fence_async_global();
      __threadfence();
      red_release_gpu_add(pend, 1);
    };
    if (g.stg) {
      // LSU path: slabs leave through coalesced st.global issued by this warp; the slot is free as soon as it has been read,
      // and a tile is published with an ordinary fence + release (no bulk-group bookkeeping)
      for (int tile = unit; tile < g.total_tiles; tile += nunits) {
        const TsTile ti = ts_decode(g, tile);
        const TsTask& t = g.t[ti.ti];
        const int m0 = ti.r * (TC_BLOCK_M * CG) + (int)rank * TC_BLOCK_M, n0 = ti.n0;
        const int nsl = ts_num_slabs(t, n0), sc = ts_slab_cols(t.kind), kind = t.kind;
        const int rows_valid = t.M - m0;
        for (int s = 0; s < nsl; ++s, ++j) {
          const int slot = (int)(j % SLOTS);
          mbar_wait(sready_bar(slot), (j / SLOTS) & 1u);
          const uint32_t src = slot_base + slot * TS_SLOT_BYTES;
          const int c0 = n0 + s * sc, cols_valid = t.N - c0;
          uint8_t* g0 = reinterpret_cast<uint8_t*>(t.o0) + ((size_t)m0 * t.N + c0) * 4;
          uint8_t* g1 = reinterpret_cast<uint8_t*>(t.o1) + ((size_t)m0 * t.N + c0) * 2;
          if (kind == TS_GRAD) {
            ts_stg_region<8>(src, g0, (size_t)t.N * 4, rows_valid, cols_valid * 4, lane);
            ts_stg_region<4>(src + TS_SLAB_F32, g1, (size_t)t.N * 2, rows_valid, cols_valid * 2, lane);
          } else if (kind == TS_GRAD_P) {
            ts_stg_region<4>(src, g0, (size_t)t.N * 4, rows_valid, cols_valid * 4, lane);
            ts_stg_region<2>(src + TS_SLAB_F32, g1, (size_t)t.N * 2, rows_valid, cols_valid * 2, lane);
          } else {
            ts_stg_region<4>(src + TS_SLAB_F32, g1, (size_t)t.N * 2, rows_valid, cols_valid * 2, lane);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(sfree_bar(slot));
        }
        if (persist) {
          __threadfence();
          __syncwarp();
          if (lane == 0) red_release_gpu_add(t.done + ti.r, 1);
        }
      }
    } else
    for (int tile = unit; tile < g.total_tiles; tile += nunits) {
      const TsTile ti = ts_decode(g, tile);
      const TsTask& t = g.t[ti.ti];
      const int m0 = ti.r * (TC_BLOCK_M * CG) + (int)rank * TC_BLOCK_M, n0 = ti.n0;
      const int nsl = ts_num_slabs(t, n0), sc = ts_slab_cols(t.kind), kind = t.kind;
      for (int s = 0; s < nsl; ++s, ++j) {
        const int slot = (int)(j % SLOTS);
        // Nothing to store yet: use the idle time to publish the previous tile (its stores must have completed).  This
        // also keeps publication independent of this CTA's own future progress -- a later tile of this very CTA may be
        // waiting for that counter when a phase has fewer tiles than the grid has CTA pairs.
        if (pend != nullptr && !__shfl_sync(0xffffffffu, (int)mbar_try_wait(sready_bar(slot), (j / SLOTS) & 1u), 0)) {
          if (lane == 0) {
            bulk_wait_all();
            if (prev_slot >= 0) mbar_arrive(sfree_bar(prev_slot));
            publish();
          }
          prev_slot = -1;
          pend = nullptr;
          __syncwarp();
        }
        mbar_wait(sready_bar(slot), (j / SLOTS) & 1u);
        if (lane == 0) {  // (one fixed lane: bulk async-groups are per thread)
          const uint32_t src = slot_base + slot * TS_SLOT_BYTES;
          if (g.hints & 4) {
            const uint64_t pf = l2_policy_evict_first();
            if (kind != TS_ERR) tma_store_2d_hint(&t.tmO0, src, n0 + s * sc, m0, pf);
            tma_store_2d_hint(&t.tmO1, src + TS_SLAB_F32, n0 + s * sc, m0, pf);
          } else {
            if (kind != TS_ERR) tma_store_2d(&t.tmO0, src, n0 + s * sc, m0);
            tma_store_2d(&t.tmO1, src + TS_SLAB_F32, n0 + s * sc, m0);
          }
          bulk_commit();
          if (prev_slot >= 0) {
            bulk_wait_read<1>();  // everything but the group just committed has been read out of shared memory
            mbar_arrive(sfree_bar(prev_slot));
          }
          if (pend != nullptr && since + 1 == TS_PUB_LAG) {
            bulk_wait<TS_PUB_LAG>();
            publish();
          }
        }
        if (pend != nullptr && ++since == TS_PUB_LAG) pend = nullptr;
        prev_slot = slot;
        __syncwarp();
      }
      if (persist) {
        if (pend != nullptr) {  // (the tile had fewer slabs than the lag: wait for everything)
          if (lane == 0) { bulk_wait_all(); publish(); }
          __syncwarp();
        }
        pend = t.done + ti.r;
        since = 0;
      }
    }
    if (pend != nullptr && lane == 0) { bulk_wait_all(); publish(); }
    if (lane == 0) bulk_wait_all();  // all stores complete (written) before the CTA may exit
  } else {
    // ===================== epilogue warps: this CTA's 128 rows, lane = row =====================
    const int ew = warp - 4, team = ew >> 2, q = warp & 3;
    const uint32_t row = (uint32_t)(q * 32 + lane);
    int as = 0;
    uint32_t aphase = 0, j = 0;
    for (int tile = unit; tile < g.total_tiles; tile += nunits) {
      const TsTile ti = ts_decode(g, tile);
      const TsTask& t = g.t[ti.ti];
      const int n0 = ti.n0, kind = t.kind;
      const int nsl = ts_num_slabs(t, n0);
      const float alpha = t.alpha, cg = (persist && ti.step == g.nsteps - 1) ? t.cg_last : t.cg;
      float* const red = t.red;
      const int m0 = ti.r * (TC_BLOCK_M * CG) + (int)rank * TC_BLOCK_M;
      const bool row_ok = m0 + (int)row < t.M;
      const int ncol_tile = t.N - n0;
      float r = 0.f;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      for (int s = 0; s < nsl; ++s, ++j) {
        if ((int)(j & 1u) != team) continue;
        const int slot = (int)(j % SLOTS);
        const uint32_t sb = slot_base + slot * TS_SLOT_BYTES;
        if (kind == TS_GRAD_P) {
          uint32_t acc[16];
          tmem_ld16(t_row + (uint32_t)(s * 16), acc);
          mbar_wait(sfull_bar(slot), (j / SLOTS) & 1u);
          tmem_ld_wait();
          ts_slab_grad_p<ACT>(acc, sb, row, alpha, cg);
        } else {
          uint32_t acc[32];
          tmem_ld32(t_row + (uint32_t)(s * TS_SLAB_COLS), acc);
          mbar_wait(sfull_bar(slot), (j / SLOTS) & 1u);
          tmem_ld_wait();
          if (kind == TS_GRAD) ts_slab_grad<ACT>(acc, sb, row, alpha, cg);
          else {
            const float rr = ts_slab_err(acc, sb, row, alpha, red != nullptr, ncol_tile - s * TS_SLAB_COLS,
                                         t.bias ? t.bias + n0 + s * TS_SLAB_COLS : nullptr);
            r += row_ok ? rr : 0.f;
          }
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
