// call.cu -- the POSITIONAL form of the C-ABI: jpcb_call(entry, desc, args[], rets[], workspace, stream).
//
// An XLA custom call (jax.ffi, what a jpc maintainer would bind -- INTEGRATION.md) receives its operands and results as two
// flat positional buffer lists plus attributes; every structured entry point of include/jpc_b200.h takes several pointer
// LISTS (weights, biases, activities, gradients) whose lengths follow from the descriptor.  This file is the one place that
// knows how a flat list splits into those lists, for every entry point, so that the XLA shim (xla_ffi_shim.cc, which cannot
// be compiled in an image without jaxlib) is a dozen lines of buffer unpacking over jpcb_call and everything else is built
// and tested here (tests/test_gpu_call.py: positional == structured, entry by entry).  No arithmetic: a call whose result
// activities are not the operand buffers themselves (no donation) copies them device-to-device first, then runs in place.
#include <cuda_runtime.h>

#include <vector>

#include "host.cuh"

using namespace jpcb;

namespace {

struct Cursor {
  const void* const* a; int na, ia = 0;
  void* const* r; int nr, ir = 0;
  template <typename T> std::vector<const T*> in(int n) {
    std::vector<const T*> v(n);
    for (int i = 0; i < n; ++i) v[i] = static_cast<const T*>(a[ia++]);
    return v;
  }
  const float* in1() { return static_cast<const float*>(a[ia++]); }
  std::vector<float*> out(int n) {
    std::vector<float*> v(n);
    for (int i = 0; i < n; ++i) v[i] = static_cast<float*>(r[ir++]);
    return v;
  }
  float* out1() { return static_cast<float*>(r[ir++]); }
};

template <typename T> const T* const* lst(const std::vector<const T*>& v) { return v.empty() ? nullptr : v.data(); }
float* const* lst(const std::vector<float*>& v) { return v.empty() ? nullptr : v.data(); }

// shapes of the PC / ePC state list: L entries z_1..z_L, or L + 1 (z_0 first) when free_input
struct PcShape {
  int L, nb, fi, La, Bl;
  const int32_t* D;
  explicit PcShape(const jpcb_pc_desc* d) : L(d->n_layers), nb(d->has_bias ? d->n_layers : 0), fi(d->free_input ? 1 : 0),
                                           La(d->n_layers + (d->free_input ? 1 : 0)), Bl(d->batch_local), D(d->dims) {}
  size_t act_elems(int i) const { return (size_t)Bl * D[fi ? i : i + 1]; }
};
struct BpcShape {
  int n, nbW, nbV, fi, Bl;
  const int32_t* D;
  explicit BpcShape(const jpcb_bpc_desc* d) : n(d->td.n_layers), nbW(d->td.has_bias ? d->td.n_layers : 0),
                                             nbV(d->bu_has_bias ? d->td.n_layers : 0), fi(d->td.free_input ? 1 : 0),
                                             Bl(d->td.batch_local), D(d->td.dims) {}
  size_t act_elems(int k) const { return (size_t)Bl * D[k + 1]; }
};

bool layers_ok(int L) { return L >= 1 && L <= JPCB_MAX_LAYERS; }

int counts(int entry, const void* desc, size_t desc_bytes, int* na, int* nr) {
  const bool bpc = entry >= JPCB_CALL_BPC_ENERGY && entry <= JPCB_CALL_BPC_TRAIN_STEP;
  if (!desc || desc_bytes != (bpc ? sizeof(jpcb_bpc_desc) : sizeof(jpcb_pc_desc)))
    return fail(JPCB_EINVAL, "jpcb_call: descriptor of %zu bytes, expected %zu", desc_bytes, bpc ? sizeof(jpcb_bpc_desc) : sizeof(jpcb_pc_desc));
  if (bpc) {
    const BpcShape s(static_cast<const jpcb_bpc_desc*>(desc));
    if (!layers_ok(s.n)) return fail(JPCB_EINVAL, "jpcb_call: n_layers = %d", s.n);
    const int in = 2 * s.n + s.nbW + s.nbV + s.n + (s.fi ? 0 : 1) + 1;   // W, bW, V, bV, A, [x], y
    const int g = 2 * s.n + s.nbW + s.nbV;
    switch (entry) {
      case JPCB_CALL_BPC_ENERGY: *na = in; *nr = 1 + 1; return JPCB_OK;
      case JPCB_CALL_BPC_ACTIVITY_GRAD: *na = in; *nr = 1 + s.n + 1; return JPCB_OK;
      case JPCB_CALL_BPC_INFER: *na = in; *nr = s.n + 1; return JPCB_OK;
      case JPCB_CALL_BPC_PARAM_GRADS: *na = in; *nr = g + 1 + 1; return JPCB_OK;
      case JPCB_CALL_BPC_TRAIN_STEP: *na = in; *nr = s.n + g + 1 + 1; return JPCB_OK;
    }
  }
  const PcShape s(static_cast<const jpcb_pc_desc*>(desc));
  if (!layers_ok(s.L)) return fail(JPCB_EINVAL, "jpcb_call: n_layers = %d", s.L);
  const int wb = s.L + s.nb, st = wb + s.La + (s.fi ? 0 : 1) + 1;           // W, b, state list, [x], y
  switch (entry) {
    case JPCB_CALL_PC_FFWD_INIT: *na = wb + 2; *nr = s.L + 1 + 1; return JPCB_OK;
    case JPCB_CALL_PC_ENERGY: *na = st; *nr = 1 + 1; return JPCB_OK;
    case JPCB_CALL_PC_ACTIVITY_GRAD: *na = st; *nr = 1 + s.La + 1; return JPCB_OK;
    case JPCB_CALL_PC_INFER: *na = st; *nr = s.La + 1; return JPCB_OK;
    case JPCB_CALL_PC_PARAM_GRADS: *na = st; *nr = wb + 1 + 1; return JPCB_OK;
    case JPCB_CALL_PC_TRAIN_STEP: *na = wb + 2; *nr = s.L + 2 + wb + 1; return JPCB_OK;
    case JPCB_CALL_HPC_PARAM_GRADS: *na = wb + 2 * s.L; *nr = wb + 1 + 1; return JPCB_OK;
    case JPCB_CALL_EPC_GRADS: *na = st; *nr = 1 + s.La + wb + 1; return JPCB_OK;
  }
  return fail(JPCB_EINVAL, "jpcb_call: unknown entry %d", entry);
}

// results that are states advanced in place: start from the operand's contents unless the caller donated the buffer
template <typename Shape>
int seed_states(const Shape& s, const std::vector<const float*>& src, const std::vector<float*>& dst, cudaStream_t st) {
  for (size_t i = 0; i < dst.size(); ++i)
    if (dst[i] != src[i]) CK(cudaMemcpyAsync(dst[i], src[i], s.act_elems((int)i) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return JPCB_OK;
}

}  // namespace

extern "C" int jpcb_call_counts(int entry, const void* desc, size_t desc_bytes, int* n_args, int* n_rets) {
  int na = 0, nr = 0;
  RET(counts(entry, desc, desc_bytes, &na, &nr));
  if (n_args) *n_args = na;
  if (n_rets) *n_rets = nr;
  return JPCB_OK;
}

extern "C" int jpcb_call(int entry, const void* desc, size_t desc_bytes, const void* const* args, int n_args,
                         void* const* rets, int n_rets, size_t workspace_bytes, void* stream) {
  int na = 0, nr = 0;
  RET(counts(entry, desc, desc_bytes, &na, &nr));
  if (n_args != na || n_rets != nr)
    return fail(JPCB_EINVAL, "jpcb_call(entry %d): %d operands / %d results, the descriptor implies %d / %d", entry, n_args, n_rets, na, nr);
  if (!args || !rets) return fail(JPCB_EINVAL, "jpcb_call: null operand / result list");
  Cursor c{args, n_args, 0, rets, n_rets, 0};
  void* ws = rets[n_rets - 1];
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  if (entry >= JPCB_CALL_BPC_ENERGY && entry <= JPCB_CALL_BPC_TRAIN_STEP) {
    const jpcb_bpc_desc* d = static_cast<const jpcb_bpc_desc*>(desc);
    const BpcShape s(d);
    auto W = c.in<float>(s.n), bW = c.in<float>(s.nbW), V = c.in<float>(s.n), bV = c.in<float>(s.nbV), A = c.in<float>(s.n);
    const float* x = s.fi ? nullptr : c.in1();
    const float* y = c.in1();
    switch (entry) {
      case JPCB_CALL_BPC_ENERGY:
        return jpcb_bpc_energy(d, lst(W), lst(bW), lst(V), lst(bV), lst(A), x, y, c.out1(), ws, workspace_bytes, stream);
      case JPCB_CALL_BPC_ACTIVITY_GRAD: {
        float* F = c.out1();
        auto gA = c.out(s.n);
        return jpcb_bpc_activity_grad(d, lst(W), lst(bW), lst(V), lst(bV), lst(A), x, y, F, lst(gA), ws, workspace_bytes, stream);
      }
      case JPCB_CALL_BPC_INFER: {
        auto A2 = c.out(s.n);
        RET(seed_states(s, A, A2, st));
        return jpcb_bpc_infer(d, lst(W), lst(bW), lst(V), lst(bV), lst(A2), x, y, ws, workspace_bytes, stream);
      }
      case JPCB_CALL_BPC_PARAM_GRADS: {
        auto gW = c.out(s.n), gbW = c.out(s.nbW), gV = c.out(s.n), gbV = c.out(s.nbV);
        return jpcb_bpc_param_grads(d, lst(W), lst(bW), lst(V), lst(bV), lst(A), x, y, lst(gW), lst(gbW), lst(gV), lst(gbV), c.out1(), ws,
                                    workspace_bytes, stream);
      }
      default: {  // JPCB_CALL_BPC_TRAIN_STEP
        auto A2 = c.out(s.n);
        auto gW = c.out(s.n), gbW = c.out(s.nbW), gV = c.out(s.n), gbV = c.out(s.nbV);
        RET(seed_states(s, A, A2, st));
        return jpcb_bpc_train_step(d, lst(W), lst(bW), lst(V), lst(bV), lst(A2), x, y, lst(gW), lst(gbW), lst(gV), lst(gbV), c.out1(), ws,
                                   workspace_bytes, stream);
      }
    }
  }

  const jpcb_pc_desc* d = static_cast<const jpcb_pc_desc*>(desc);
  const PcShape s(d);
  auto W = c.in<float>(s.L), b = c.in<float>(s.nb);
  switch (entry) {
    case JPCB_CALL_PC_FFWD_INIT: {
      const float* x = c.in1();
      const float* y = c.in1();
      auto A = c.out(s.L);
      return jpcb_pc_ffwd_init(d, lst(W), lst(b), x, lst(A), y, c.out1(), ws, workspace_bytes, stream);
    }
    case JPCB_CALL_PC_TRAIN_STEP: {
      const float* x = c.in1();
      const float* y = c.in1();
      auto A = c.out(s.L);
      float* loss = c.out1();
      float* E = c.out1();
      auto gW = c.out(s.L), gb = c.out(s.nb);
      return jpcb_pc_train_step(d, lst(W), lst(b), x, y, lst(A), loss, E, lst(gW), lst(gb), ws, workspace_bytes, stream);
    }
    case JPCB_CALL_HPC_PARAM_GRADS: {
      auto U = c.in<float>(s.L), T = c.in<float>(s.L);
      auto gW = c.out(s.L), gb = c.out(s.nb);
      return jpcb_hpc_param_grads(d, lst(W), lst(b), lst(U), lst(T), lst(gW), lst(gb), c.out1(), ws, workspace_bytes, stream);
    }
    default: break;
  }
  auto A = c.in<float>(s.La);
  const float* x = s.fi ? nullptr : c.in1();
  const float* y = c.in1();
  switch (entry) {
    case JPCB_CALL_PC_ENERGY:
      return jpcb_pc_energy(d, lst(W), lst(b), lst(A), x, y, c.out1(), ws, workspace_bytes, stream);
    case JPCB_CALL_PC_ACTIVITY_GRAD: {
      float* F = c.out1();
      auto gA = c.out(s.La);
      return jpcb_pc_activity_grad(d, lst(W), lst(b), lst(A), x, y, F, lst(gA), ws, workspace_bytes, stream);
    }
    case JPCB_CALL_PC_INFER: {
      auto A2 = c.out(s.La);
      RET(seed_states(s, A, A2, st));
      return jpcb_pc_infer(d, lst(W), lst(b), lst(A2), x, y, nullptr, ws, workspace_bytes, stream);
    }
    case JPCB_CALL_PC_PARAM_GRADS: {
      auto gW = c.out(s.L), gb = c.out(s.nb);
      return jpcb_pc_param_grads(d, lst(W), lst(b), lst(A), x, y, lst(gW), lst(gb), c.out1(), ws, workspace_bytes, stream);
    }
    default: {  // JPCB_CALL_EPC_GRADS
      float* F = c.out1();
      auto gE = c.out(s.La);
      auto gW = c.out(s.L), gb = c.out(s.nb);
      return jpcb_epc_grads(d, lst(W), lst(b), lst(A), x, y, F, lst(gE), lst(gW), lst(gb), ws, workspace_bytes, stream);
    }
  }
}
