// xla_ffi_shim.cc -- jax.ffi (XLA custom-call) registration of the C-ABI in include/jpc_b200.h.
//
// NOT compiled in this image: it needs jaxlib's headers (jaxlib/include/xla/ffi/api/ffi.h), and jax / jaxlib are not
// installed here.  jpc_b200/_build.py compiles it into libjpc_b200_xla.so only when `import jaxlib` succeeds
// and the header is found.  To keep the part that cannot be built here as small as possible, everything that depends on the
// entry point -- how the flat operand / result lists split into weight, bias, state and gradient lists, the operand counts
// a descriptor implies, the device copy of non-donated states -- lives in the library proper (csrc/call.cu: jpcb_call /
// jpcb_call_counts), where it IS built and tested (tests/test_gpu_call.py, tests/test_abi.py).  What is left here is one
// generic handler: unpack XLA's buffers into two pointer arrays, pass the descriptor attribute through, enqueue on XLA's
// stream.  One handler symbol per JPCB_CALL_* entry (tests/test_abi.py checks this list against the header's enum).
// tests/test_xla_shim_mock.py compiles THIS file (-Wall -Wextra -Werror) against a mock of the FFI header that restates the
// API surface used below and drives all 13 handlers with fake buffers and a recording jpcb_call: well-formed for that API
// shape and marshalling verified, though never against the real header.
// INTEGRATION.md shows the Python side (jax.ffi.register_ffi_target / jax.ffi.ffi_call) a jpc maintainer would add.
#include <cstdint>
#include <vector>

#include <cuda_runtime_api.h>  // cudaStream_t

#include "xla/ffi/api/ffi.h"

#include "../../include/jpc_b200.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error to_error(int code) {
  if (code == JPCB_OK) return ffi::Error::Success();
  const ffi::ErrorCode ec = (code == JPCB_EINVAL || code == JPCB_ENOTIMPL) ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal;
  return ffi::Error(ec, jpcb_last_error());
}

// The descriptor travels as one opaque byte-array attribute (static per jit trace, like every non-array argument of the
// reference's filter_jit'ed functions, jpc/_train.py:35); jpcb_call checks its size against the entry's descriptor type.
template <int ENTRY>
ffi::Error Call(cudaStream_t stream, ffi::Span<const uint8_t> desc, ffi::RemainingArgs args, ffi::RemainingRets rets) {
  if (rets.size() == 0) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "jpc_b200: the last result must be the workspace");
  std::vector<const void*> a(args.size());
  std::vector<void*> r(rets.size());
  for (size_t i = 0; i < args.size(); ++i) {
    auto buf = args.get<ffi::AnyBuffer>(i);
    if (!buf.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "jpc_b200: operand is not a buffer");
    a[i] = buf->untyped_data();
  }
  size_t ws_bytes = 0;
  for (size_t i = 0; i < rets.size(); ++i) {
    auto buf = rets.get<ffi::AnyBuffer>(i);
    if (!buf.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "jpc_b200: result is not a buffer");
    r[i] = (*buf)->untyped_data();
    if (i + 1 == rets.size()) ws_bytes = (*buf)->size_bytes();
  }
  return to_error(jpcb_call(ENTRY, desc.begin(), desc.size(), a.data(), (int)a.size(), r.data(), (int)r.size(), ws_bytes, stream));
}

}  // namespace

#define JPCB_XLA_HANDLER(symbol, entry)                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, Call<entry>,                                              \
                                ffi::Ffi::Bind()                                                  \
                                    .Ctx<ffi::PlatformStream<cudaStream_t>>()                     \
                                    .Attr<ffi::Span<const uint8_t>>("desc")                       \
                                    .RemainingArgs()                                              \
                                    .RemainingRets())

JPCB_XLA_HANDLER(jpcb_xla_pc_ffwd_init, JPCB_CALL_PC_FFWD_INIT);
JPCB_XLA_HANDLER(jpcb_xla_pc_energy, JPCB_CALL_PC_ENERGY);
JPCB_XLA_HANDLER(jpcb_xla_pc_activity_grad, JPCB_CALL_PC_ACTIVITY_GRAD);
JPCB_XLA_HANDLER(jpcb_xla_pc_infer, JPCB_CALL_PC_INFER);
JPCB_XLA_HANDLER(jpcb_xla_pc_param_grads, JPCB_CALL_PC_PARAM_GRADS);
JPCB_XLA_HANDLER(jpcb_xla_pc_train_step, JPCB_CALL_PC_TRAIN_STEP);
JPCB_XLA_HANDLER(jpcb_xla_hpc_param_grads, JPCB_CALL_HPC_PARAM_GRADS);
JPCB_XLA_HANDLER(jpcb_xla_epc_grads, JPCB_CALL_EPC_GRADS);
JPCB_XLA_HANDLER(jpcb_xla_bpc_energy, JPCB_CALL_BPC_ENERGY);
JPCB_XLA_HANDLER(jpcb_xla_bpc_activity_grad, JPCB_CALL_BPC_ACTIVITY_GRAD);
JPCB_XLA_HANDLER(jpcb_xla_bpc_infer, JPCB_CALL_BPC_INFER);
JPCB_XLA_HANDLER(jpcb_xla_bpc_param_grads, JPCB_CALL_BPC_PARAM_GRADS);
JPCB_XLA_HANDLER(jpcb_xla_bpc_train_step, JPCB_CALL_BPC_TRAIN_STEP);
