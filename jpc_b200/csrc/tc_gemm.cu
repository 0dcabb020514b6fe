// tc_gemm.cu -- host side of the tcgen05 grouped GEMM (tc_gemm.cuh): TMA tensor maps, tile counts,
// cluster launch, and the choice of the compile-time epilogue specialisation for a launch.
#include <cuda.h>

#include <atomic>

#include "host.cuh"
#include "tc_gemm.cuh"
#include "tc_staged.cuh"

namespace jpcb {
namespace {

typedef __nv_bfloat16 bf16;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return (EncodeTiledFn) nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// bf16 [rows, cols] row-major, box = {64 cols, box_rows}, SWIZZLE_128B, OOB reads give zeros
// (K-major operand: rows = M/N, cols = K; MN-major operand: rows = K, cols = M/N, box_rows = 64)
int make_map(CUtensorMap* m, const bf16* p, int rows, int cols, int box_rows) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return fail(JPCB_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(JPCB_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d", (int)r, rows, cols);
  return JPCB_OK;
}

// epilogue slab of the staged kernel (tc_staged.cuh): [rows, cols] row-major fp32 or bf16, box = {32 columns, 128 rows};
// the swizzle span equals the box's row bytes (128 B fp32 / 64 B bf16) so that row-owner threads read conflict-free
int make_slab_map(CUtensorMap* m, const void* p, int rows, int cols, bool is_f32) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return fail(JPCB_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const int es = is_f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * es};
  cuuint32_t box[2] = {(cuuint32_t)TS_SLAB_COLS, (cuuint32_t)TC_BLOCK_M};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, is_f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(JPCB_ECUDA, "cuTensorMapEncodeTiled (slab) failed (%d) rows=%d cols=%d", (int)r, rows, cols);
  return JPCB_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE function attribute: remember per device (bit d of `done`)
// that it has been set for this kernel.  Setting it twice is harmless, so a relaxed race between two host threads is too.
template <typename Kern>
int ensure_dyn_smem(Kern kernel, int bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  const unsigned long long bit = dev < 64 ? (1ull << dev) : 0ull;
  if (bit && (done.load(std::memory_order_relaxed) & bit)) return JPCB_OK;
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (bit) done.fetch_or(bit, std::memory_order_relaxed);
  return JPCB_OK;
}

// cluster of two CTAs + programmatic dependent launch (the kernels execute griddepcontrol.wait before their first
// global-memory access)
template <typename Kern, typename Group>
int launch_pair_kernel(Kern kernel, const Group& g, int grid, int threads, int smem, cudaStream_t st, int cluster) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (tc_flags() & 16) ? 1 : 2;
  CK(cudaLaunchKernelEx(&cfg, kernel, g));
  count_launch();
  return JPCB_OK;
}

template <int CG, int BN, int MODE, int SUB, int ACT, int ELSRC, int PLAIN, bool MNMAJ = false, int EW = TC_EPI_WARPS>
int launch_variant(const TcGroup& g, int grid, cudaStream_t st) {
  using Cfg = TcCfg<CG, BN>;
  constexpr int SMEM = Cfg::smem_bytes(EW);
  static_assert(SMEM <= 227 * 1024, "shared memory of the tensor-core kernel exceeds the per-CTA limit");
  static std::atomic<unsigned long long> attr_done{0};
  RET(ensure_dyn_smem(tc_grouped_gemm<CG, MODE, SUB, ACT, ELSRC, PLAIN, MNMAJ, EW, BN>, SMEM, attr_done));
  return launch_pair_kernel(tc_grouped_gemm<CG, MODE, SUB, ACT, ELSRC, PLAIN, MNMAJ, EW, BN>, g, grid, TcEpi<EW>::THREADS, SMEM, st, CG);
}

// ---- staged-epilogue kernel (tc_staged.cuh) ----
int staged_flags() {  // JPCB_STAGED bit 0: GRAD (Euler, plain) phases, bit 1: ERR (plain) phases
  static const int f = [] { const char* v = getenv("JPCB_STAGED"); return v ? atoi(v) : 3; }();
  return f;
}
// 0: not eligible; else the activation the kernel is compiled for (ERR tasks join either)
int staged_kind(const GTask& t, int key) {
  const Epi& e = t.e;
  if (t.split || t.mn_major || e.ob_part != 0 || !e.Ob) return -1;
  if ((key == 3 || key == 5) && (staged_flags() & 1) && e.O && !e.Oh && e.Elb && e.Z) return TS_GRAD;
  if (key == 1 && (staged_flags() & 2) && !e.O && e.T) return TS_ERR;
  return -1;
}
int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
template <int ACT, int STAGES, int SLOTS>
int launch_staged_cfg(const TsGroup& g, int grid, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_done{0};
  constexpr int SMEM = ts_smem(STAGES, SLOTS);
  static_assert(SMEM <= 227 * 1024, "staged kernel: shared memory over the per-CTA limit");
  RET(ensure_dyn_smem(tc_staged_gemm<ACT, STAGES, SLOTS>, SMEM, attr_done));
  return launch_pair_kernel(tc_staged_gemm<ACT, STAGES, SLOTS>, g, grid, TS_THREADS, SMEM, st, 2);
}
template <int ACT>
int launch_staged_act(TsGroup& g, int grid, cudaStream_t st) {
  static const int cfg = env_int("JPCB_TS_CFG", 0), pf = env_int("JPCB_TS_PF", 0), hints = env_int("JPCB_TS_HINTS", 0);
  static const int dbg = env_int("JPCB_TS_DBG", 0);
  g.pf_dist = pf; g.hints = hints; g.dbg = dbg;
  if (cfg == 1) return launch_staged_cfg<ACT, 5, 2>(g, grid, st);
  if (cfg == 2) return launch_staged_cfg<ACT, 3, 5>(g, grid, st);
  return launch_staged_cfg<ACT, TS_STAGES, TS_SLOTS>(g, grid, st);
}
int fill_staged_task(TsTask& s, const GTask& t, int kind, int tile_begin) {
  const Epi& e = t.e;
  memset(&s, 0, sizeof(s));
  RET(make_map(&s.tmA, t.a_b, e.M, t.K, TC_BLOCK_M));
  RET(make_map(&s.tmB, t.b_b, e.N, t.K, TC_BLOCK_N / 2));
  if (kind == TS_GRAD) {
    RET(make_slab_map(&s.tmI0, e.Z, e.M, e.N, true));
    RET(make_slab_map(&s.tmI1, e.Elb, e.M, e.N, false));
    RET(make_slab_map(&s.tmO0, e.O, e.M, e.N, true));
  } else {
    RET(make_slab_map(&s.tmI0, e.T, e.M, e.N, true));
  }
  RET(make_slab_map(&s.tmO1, e.Ob, e.M, e.N, false));
  s.K = t.K; s.tile_begin = tile_begin; s.tiles_n = (e.N + TC_BLOCK_N - 1) / TC_BLOCK_N; s.kind = kind;
  s.M = e.M; s.N = e.N; s.alpha = e.alpha; s.cg = e.c * e.gscale; s.red = kind == TS_ERR ? e.red : nullptr;
  return JPCB_OK;
}

// Epilogue specialisations compiled in: what the timed configurations run (errors, tanh / relu
// Euler updates with e_l from the bf16 error copy or from the hoisted layer-0 prediction, scaled
// weight-gradient outputs).  Anything else takes the generic descriptor-driven kernel.
int variant_key(const GTask& t) {
  const Epi& e = t.e;
  if (t.mn_major) return 11;  // (only the weight-gradient SCALE tasks use MN-major operands)
  if (tc_flags() & 4) return 0;
  const bool noextra = e.skip == 0.f && e.bias == nullptr && e.escale == 1.f && e.ad == 0.f;
  const bool plain = noextra && e.Gin == nullptr && e.gscale == 1.f;
  if (e.mode == EPI_SCALE) return 2;
  if (t.split) {
    if (e.mode == EPI_ERR) return 20;

    // fp32-parity path: the specialised variants use fast activation approximations, so only activations that are exact
    // either way (relu, identity) may take them; e_l comes from the fp32 error buffers (ELSRC == 2)
    if (e.mode == EPI_FFWD && (e.act == ACT_RELU || e.act == ACT_LINEAR)) return e.act == ACT_RELU ? 21 : 22;
    if (e.mode == EPI_GRAD && noextra && e.act == ACT_RELU && e.El) {
      if (e.sub == GRAD_OUT && e.Gin == nullptr) return 17;
      if (e.sub == GRAD_EULER && e.Gin != nullptr) return 18;
      if (e.sub == GRAD_EULER && e.gscale == 1.f) return 19;
    }
    return 0;
  }
  if (e.mode == EPI_ERR) return plain ? 1 : 7;
  const bool act_ok = e.act == ACT_TANH || e.act == ACT_RELU;
  if (e.mode == EPI_GRAD && e.sub == GRAD_EULER && plain && act_ok && !e.El && (e.Elb || e.P))
    return 3 + (e.act == ACT_RELU ? 2 : 0) + (e.Elb ? 0 : 1);
  // bPC: forward-direction half (GRAD_OUT into scratch) and backward-direction half + combination (Gin) + Euler update
  if (e.mode == EPI_GRAD && noextra && act_ok && !e.El && e.Elb) {
    if (e.sub == GRAD_OUT && e.Gin == nullptr) return e.act == ACT_TANH ? 12 : 13;
    if (e.sub == GRAD_EULER && e.Gin != nullptr) return e.act == ACT_TANH ? 14 : 15;
  }
  if (e.mode == EPI_FFWD && (act_ok || e.act == ACT_LINEAR)) return e.act == ACT_TANH ? 8 : (e.act == ACT_RELU ? 9 : 10);
  return 0;
}

template <int CG, int BN>
int launch_group(const TcGroup& g, int key, int grid, cudaStream_t st) {
  switch (key) {
    case 1: return launch_variant<CG, BN, EPI_ERR, 0, 0, -1, 1>(g, grid, st);
    case 7: return launch_variant<CG, BN, EPI_ERR, 0, 0, -1, 0>(g, grid, st);
    case 2: return launch_variant<CG, BN, EPI_SCALE, 0, 0, -1, 0>(g, grid, st);
    // the Euler update with e_l from the bf16 error copy -- the phase that is bound by its epilogue's memory latency, not by
    // the MMAs -- runs 12 epilogue warps (3 per scheduler instead of 2; 126 registers, 2-deep load ring): 207 -> 198 us on c4
    case 3: return launch_variant<CG, BN, EPI_GRAD, GRAD_EULER, ACT_TANH, 0, 1, false, BN == 256 ? 12 : 8>(g, grid, st);
    case 4: return launch_variant<CG, BN, EPI_GRAD, GRAD_EULER, ACT_TANH, 1, 1>(g, grid, st);
    case 5: return launch_variant<CG, BN, EPI_GRAD, GRAD_EULER, ACT_RELU, 0, 1, false, BN == 256 ? 12 : 8>(g, grid, st);
    case 6: return launch_variant<CG, BN, EPI_GRAD, GRAD_EULER, ACT_RELU, 1, 1>(g, grid, st);
    case 8: return launch_variant<CG, BN, EPI_FFWD, 0, ACT_TANH, -1, 0>(g, grid, st);
    case 9: return launch_variant<CG, BN, EPI_FFWD, 0, ACT_RELU, -1, 0>(g, grid, st);
    case 10: return launch_variant<CG, BN, EPI_FFWD, 0, ACT_LINEAR, -1, 0>(g, grid, st);
    case 11: return launch_variant<CG, BN, EPI_SCALE, 0, 0, -1, 0, true>(g, grid, st);
    case 12: return launch_variant<CG, BN, EPI_GRAD, GRAD_OUT, ACT_TANH, 0, 1>(g, grid, st);
    case 13: return launch_variant<CG, BN, EPI_GRAD, GRAD_OUT, ACT_RELU, 0, 1>(g, grid, st);
    case 14: return launch_variant<CG, BN, EPI_GRAD, GRAD_EULER, ACT_TANH, 0, 2>(g, grid, st);
    case 15: return launch_variant<CG, BN, EPI_GRAD, GRAD_EULER, ACT_RELU, 0, 2>(g, grid, st);
    case 17: return launch_variant<CG, BN, EPI_GRAD, GRAD_OUT, ACT_RELU, 2, 1>(g, grid, st);
    case 18: return launch_variant<CG, BN, EPI_GRAD, GRAD_EULER, ACT_RELU, 2, 2>(g, grid, st);
    case 19: return launch_variant<CG, BN, EPI_GRAD, GRAD_EULER, ACT_RELU, 2, 1>(g, grid, st);
    // split-capable twins of the ERR / FFWD variants (ELSRC == 2 marks "may write 3-way split operand copies")
    case 20: return launch_variant<CG, BN, EPI_ERR, 0, 0, 2, 0>(g, grid, st);
    case 21: return launch_variant<CG, BN, EPI_FFWD, 0, ACT_RELU, 2, 0>(g, grid, st);
    case 22: return launch_variant<CG, BN, EPI_FFWD, 0, ACT_LINEAR, 2, 0>(g, grid, st);
    default: return launch_variant<CG, BN, -1, -1, -1, -1, 0>(g, grid, st);
  }
}

template <int CG>
int launch_tc_cg(const std::vector<GTask>& tasks, cudaStream_t st) {
  using Cfg = TcCfg<CG>;
  constexpr int BM = TC_BLOCK_M * CG;  // rows of one tile (owned by a CTA, or by a CTA pair)
  // tasks of one phase are independent: launch them grouped by epilogue specialisation
  std::vector<int> keys(tasks.size());
  std::vector<char> done(tasks.size(), 0);
  for (size_t i = 0; i < tasks.size(); ++i) {
    if (tasks[i].mn_major && tasks[i].e.mode != EPI_SCALE) return fail(JPCB_EINVAL, "internal: MN-major operands are only built for SCALE tasks");
    keys[i] = variant_key(tasks[i]);
  }
  for (size_t first = 0; first < tasks.size(); ++first) {
    if (done[first]) continue;
    const int key = keys[first];
    // tile width of this group: 256, or 128 when that needs fewer (half-weight) rounds of the persistent grid -- launches
    // with only one or two waves of 256-wide tiles lose up to half a wave to quantisation
    const int units = num_sms() / CG;  // persistent: one CTA (pair) per SM (pair)
    int BN = TC_BLOCK_N;
    {
      long long t256 = 0, t128 = 0;
      for (size_t j = first; j < tasks.size(); ++j) {
        if (done[j] || keys[j] != key) continue;
        const long long tm = (tasks[j].e.M + BM - 1) / BM;
        t256 += tm * ((tasks[j].e.N + 255) / 256);
        t128 += tm * ((tasks[j].e.N + 127) / 128);
      }
      const long long r256 = 2 * ((t256 + units - 1) / units), r128 = (t128 + units - 1) / units;  // in half-tile rounds
      static const int allow128 = [] { const char* v = getenv("JPCB_TC_BN128"); return v ? atoi(v) : 1; }();
      if (allow128 && r128 < r256 && t256 <= 4LL * units) BN = 128;
    }
    size_t i = first;
    // staged-epilogue kernel (tc_staged.cuh): the plain ERR / Euler-GRAD phases of wide layers, 256-wide tiles only
    if (CG == 2 && BN == TC_BLOCK_N && staged_kind(tasks[first], key) >= 0) {
      bool all_ok = true;
      for (size_t j = first; j < tasks.size(); ++j)
        if (!done[j] && keys[j] == key && staged_kind(tasks[j], key) < 0) all_ok = false;
      if (all_ok) {
        while (i < tasks.size()) {
          // (the parameter block is large: keep it off the stack of this recursive-free function by making it static thread-local)
          static thread_local TsGroup sg;
          memset(&sg, 0, 64);
          int n = 0, tiles = 0;
          for (; i < tasks.size() && n < TS_MAX_TASKS; ++i) {
            if (done[i] || keys[i] != key) continue;
            done[i] = 1;
            const GTask& t = tasks[i];
            if (t.K <= 0 || !t.a_b || !t.b_b) return fail(JPCB_EINVAL, "internal: tensor-core task without bf16 operands");
            RET(fill_staged_task(sg.t[n], t, staged_kind(t, key), tiles));
            tiles += sg.t[n].tiles_n * ((t.e.M + BM - 1) / BM);
            ++n;
          }
          if (n == 0) break;
          sg.ntasks = n;
          sg.total_tiles = tiles;
          const int grid = CG * (tiles < units ? tiles : units);
          if (key == 5) RET(launch_staged_act<ACT_RELU>(sg, grid, st));
          else RET(launch_staged_act<ACT_TANH>(sg, grid, st));
        }
        continue;
      }
    }
    while (i < tasks.size()) {
      TcGroup g;
      memset(&g, 0, 64);
      int n = 0, tiles = 0;
      for (; i < tasks.size() && n < TC_MAX_TASKS; ++i) {
        if (done[i] || keys[i] != key) continue;
        done[i] = 1;
        const GTask& t = tasks[i];
        if (t.K <= 0 || !t.a_b || !t.b_b) return fail(JPCB_EINVAL, "internal: tensor-core task without bf16 operands");
        TcTask& s = g.t[n];
        if (t.split && (t.a_ld <= 0 || t.b_ld <= 0 || t.a_part % TC_BLOCK_K || t.b_part % TC_BLOCK_K))
          return fail(JPCB_EINVAL, "internal: split task without padded operand geometry");
        if (t.mn_major) {
          RET(make_map(&s.tmA, t.a_b, t.K, t.split ? t.a_ld : t.e.M, 64));
          RET(make_map(&s.tmB, t.b_b, t.K, t.split ? t.b_ld : t.e.N, 64));
        } else {
          RET(make_map(&s.tmA, t.a_b, t.e.M, t.split ? t.a_ld : t.K, TC_BLOCK_M));
          RET(make_map(&s.tmB, t.b_b, t.e.N, t.split ? t.b_ld : t.K, BN / CG));
        }
        s.split = t.split; s.a_part = t.a_part; s.b_part = t.b_part;
        s.pad_[0] = s.pad_[1] = 0;
        s.K = t.K;
        s.tile_begin = tiles;
        s.tiles_n = (t.e.N + BN - 1) / BN;
        s.e = t.e;
        tiles += s.tiles_n * ((t.e.M + BM - 1) / BM);
        ++n;
      }
      if (n == 0) break;
      g.ntasks = n;
      g.total_tiles = tiles;
      const int grid = CG * (tiles < units ? tiles : units);
      if (BN == 128) RET((launch_group<CG, 128>(g, key, grid, st)));
      else RET((launch_group<CG, 256>(g, key, grid, st)));
    }
  }
  return JPCB_OK;
}

}  // namespace

int launch_tc(const std::vector<GTask>& tasks, cudaStream_t st) {
  if (tasks.empty()) return JPCB_OK;
#ifdef JPCB_TC_WITH_CG1
  if (tc_flags() & 2) return launch_tc_cg<1>(tasks, st);
#endif
  return launch_tc_cg<2>(tasks, st);
}

}  // namespace jpcb
