// xla_ffi_shim.cc -- jax.ffi (XLA custom-call) registration of the C-ABI in include/jpc_b200.h.
//
// NOT compiled in this image: it needs jaxlib's headers (jaxlib/include/xla/ffi/api/ffi.h), and
// jax / jaxlib are not installed here.  jpc_b200/_build.py compiles it into
// libjpc_b200_xla.so only when `python -c "import jaxlib"` succeeds and the header is found.  It is
// the thin layer north_star asks for: each handler unpacks XLA buffers into the plain pointers of
// the C-ABI and enqueues on XLA's stream; no arithmetic lives here.  INTEGRATION.md shows the
// Python side (jax.ffi.register_ffi_target / jax.ffi.ffi_call) that a jpc maintainer would add.
#include <cstdint>
#include <cstring>
#include <vector>

#include "xla/ffi/api/ffi.h"

#include "../../include/jpc_b200.h"

namespace ffi = xla::ffi;

namespace {

// The descriptor travels as one opaque byte-array attribute (static per jit trace, like every
// non-array argument of the reference's filter_jit'ed functions, jpc/_train.py:35).
const jpcb_pc_desc* desc_of(ffi::Span<const uint8_t> bytes) {
  return bytes.size() == sizeof(jpcb_pc_desc) ? reinterpret_cast<const jpcb_pc_desc*>(bytes.begin()) : nullptr;
}

ffi::Error to_error(int code) {
  if (code == JPCB_OK) return ffi::Error::Success();
  const ffi::ErrorCode ec = (code == JPCB_EINVAL || code == JPCB_ENOTIMPL) ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal;
  return ffi::Error(ec, jpcb_last_error());
}

// args: W_0..W_{L-1}, [b_0..b_{L-1}], x, y        rets: A_0..A_{L-1}, loss, E_layers, gW_0.., [gb_0..], workspace
ffi::Error PcTrainStepImpl(cudaStream_t stream, ffi::Span<const uint8_t> desc_bytes, ffi::RemainingArgs args, ffi::RemainingRets rets) {
  const jpcb_pc_desc* d = desc_of(desc_bytes);
  if (!d) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "bad jpcb_pc_desc attribute");
  const int L = d->n_layers, nb = d->has_bias ? L : 0;
  if ((int)args.size() != L + nb + 2 || (int)rets.size() != L + 2 + L + nb + 1)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "jpcb_pc_train_step: wrong operand/result count");
  std::vector<const float*> W(L), b(L, nullptr);
  std::vector<float*> A(L), gW(L), gb(L, nullptr);
  size_t i = 0;
  for (int l = 0; l < L; ++l) W[l] = reinterpret_cast<const float*>(args.get<ffi::AnyBuffer>(i++)->untyped_data());
  for (int l = 0; l < nb; ++l) b[l] = reinterpret_cast<const float*>(args.get<ffi::AnyBuffer>(i++)->untyped_data());
  const float* x = reinterpret_cast<const float*>(args.get<ffi::AnyBuffer>(i++)->untyped_data());
  const float* y = reinterpret_cast<const float*>(args.get<ffi::AnyBuffer>(i++)->untyped_data());
  size_t r = 0;
  for (int l = 0; l < L; ++l) A[l] = reinterpret_cast<float*>((*rets.get<ffi::AnyBuffer>(r++))->untyped_data());
  float* loss = reinterpret_cast<float*>((*rets.get<ffi::AnyBuffer>(r++))->untyped_data());
  float* E = reinterpret_cast<float*>((*rets.get<ffi::AnyBuffer>(r++))->untyped_data());
  for (int l = 0; l < L; ++l) gW[l] = reinterpret_cast<float*>((*rets.get<ffi::AnyBuffer>(r++))->untyped_data());
  for (int l = 0; l < nb; ++l) gb[l] = reinterpret_cast<float*>((*rets.get<ffi::AnyBuffer>(r++))->untyped_data());
  auto ws = *rets.get<ffi::AnyBuffer>(r++);
  return to_error(jpcb_pc_train_step(d, W.data(), nb ? b.data() : nullptr, x, y, A.data(), loss, E, gW.data(), nb ? gb.data() : nullptr,
                                     ws->untyped_data(), ws->size_bytes(), stream));
}

// args: W.., [b..], A_0..A_{L-1}, x, y            rets: A'_0..A'_{L-1} (aliased to A via input_output_aliases), workspace
ffi::Error PcInferImpl(cudaStream_t stream, ffi::Span<const uint8_t> desc_bytes, ffi::RemainingArgs args, ffi::RemainingRets rets) {
  const jpcb_pc_desc* d = desc_of(desc_bytes);
  if (!d) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "bad jpcb_pc_desc attribute");
  const int L = d->n_layers, nb = d->has_bias ? L : 0;
  if ((int)args.size() != L + nb + L + 2 || (int)rets.size() != L + 1)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "jpcb_pc_infer: wrong operand/result count");
  std::vector<const float*> W(L), b(L, nullptr);
  std::vector<float*> A(L);
  size_t i = 0;
  for (int l = 0; l < L; ++l) W[l] = reinterpret_cast<const float*>(args.get<ffi::AnyBuffer>(i++)->untyped_data());
  for (int l = 0; l < nb; ++l) b[l] = reinterpret_cast<const float*>(args.get<ffi::AnyBuffer>(i++)->untyped_data());
  i += L;  // the activities are donated: results alias them, the solve runs in place on the results
  const float* x = reinterpret_cast<const float*>(args.get<ffi::AnyBuffer>(i++)->untyped_data());
  const float* y = reinterpret_cast<const float*>(args.get<ffi::AnyBuffer>(i++)->untyped_data());
  for (int l = 0; l < L; ++l) A[l] = reinterpret_cast<float*>((*rets.get<ffi::AnyBuffer>(l))->untyped_data());
  auto ws = *rets.get<ffi::AnyBuffer>(L);
  return to_error(jpcb_pc_infer(d, W.data(), nb ? b.data() : nullptr, A.data(), x, y, nullptr, ws->untyped_data(), ws->size_bytes(), stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(jpcb_xla_pc_train_step, PcTrainStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<ffi::Span<const uint8_t>>("desc")
                                  .RemainingArgs()
                                  .RemainingRets());
XLA_FFI_DEFINE_HANDLER_SYMBOL(jpcb_xla_pc_infer, PcInferImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<ffi::Span<const uint8_t>>("desc")
                                  .RemainingArgs()
                                  .RemainingRets());
