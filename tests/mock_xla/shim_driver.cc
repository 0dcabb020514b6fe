// TEST INFRASTRUCTURE -- drives every handler of jpc_b200/csrc/xla_ffi_shim.cc (compiled against the mock XLA FFI header next to
// this file) with fake buffers and a recording stand-in for jpcb_call, and checks what the shim forwards: entry id, descriptor
// bytes, operand / result pointers in order, workspace size = size of the LAST result, the stream; and that a library error
// comes back as an ffi::Error carrying jpcb_last_error().  Prints "shim ok: N handlers" and exits 0 on success.
#include <cstdio>
#include <cstring>
#include <vector>

#include "xla/ffi/api/ffi.h"

#include "../../include/jpc_b200.h"

namespace {
struct Seen {
  int entry = -1, n_args = 0, n_rets = 0, calls = 0;
  const void* desc = nullptr; size_t desc_bytes = 0, ws_bytes = 0; void* stream = nullptr;
  std::vector<const void*> args; std::vector<void*> rets;
} seen;
int next_code = JPCB_OK;
}  // namespace

extern "C" int jpcb_call(int entry, const void* desc, size_t desc_bytes, const void* const* args, int n_args, void* const* rets,
                         int n_rets, size_t workspace_bytes, void* stream) {
  seen.calls++; seen.entry = entry; seen.desc = desc; seen.desc_bytes = desc_bytes; seen.n_args = n_args; seen.n_rets = n_rets;
  seen.ws_bytes = workspace_bytes; seen.stream = stream;
  seen.args.assign(args, args + n_args); seen.rets.assign(rets, rets + n_rets);
  return next_code;
}
extern "C" const char* jpcb_last_error(void) { return "recorded failure"; }

#define DECL(sym) extern "C" int sym(xla::ffi::MockCallFrame*);
DECL(jpcb_xla_pc_ffwd_init) DECL(jpcb_xla_pc_energy) DECL(jpcb_xla_pc_activity_grad) DECL(jpcb_xla_pc_infer)
DECL(jpcb_xla_pc_param_grads) DECL(jpcb_xla_pc_train_step) DECL(jpcb_xla_hpc_param_grads) DECL(jpcb_xla_epc_grads)
DECL(jpcb_xla_bpc_energy) DECL(jpcb_xla_bpc_activity_grad) DECL(jpcb_xla_bpc_infer) DECL(jpcb_xla_bpc_param_grads)
DECL(jpcb_xla_bpc_train_step)

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s (line %d, handler %d)\n", #c, __LINE__, h); return 1; } } while (0)

int main() {
  typedef int (*Handler)(xla::ffi::MockCallFrame*);
  const Handler handlers[JPCB_CALL_N_ENTRIES] = {
      jpcb_xla_pc_ffwd_init, jpcb_xla_pc_energy, jpcb_xla_pc_activity_grad, jpcb_xla_pc_infer, jpcb_xla_pc_param_grads,
      jpcb_xla_pc_train_step, jpcb_xla_hpc_param_grads, jpcb_xla_epc_grads, jpcb_xla_bpc_energy, jpcb_xla_bpc_activity_grad,
      jpcb_xla_bpc_infer, jpcb_xla_bpc_param_grads, jpcb_xla_bpc_train_step};
  static char arena[64][8];
  for (int h = 0; h < JPCB_CALL_N_ENTRIES; ++h) {
    xla::ffi::MockCallFrame f;
    f.stream = &arena[63];
    f.desc.assign(17 + h, (uint8_t)h);
    const int na = 3 + h, nr = 2 + h % 3;
    for (int i = 0; i < na; ++i) f.args.emplace_back(&arena[i], 100 + i);
    for (int i = 0; i < nr; ++i) f.rets.emplace_back(&arena[32 + i], 1000 + i);
    next_code = JPCB_OK;
    const int before = seen.calls;
    CHECK(handlers[h](&f) == 0 && f.result.success());
    CHECK(seen.calls == before + 1 && seen.entry == h);   // handler h forwards entry id h (the header's enum order)
    CHECK(seen.desc_bytes == (size_t)(17 + h) && std::memcmp(seen.desc, f.desc.data(), f.desc.size()) == 0);
    CHECK(seen.n_args == na && seen.n_rets == nr && seen.stream == f.stream);
    for (int i = 0; i < na; ++i) CHECK(seen.args[i] == &arena[i]);
    for (int i = 0; i < nr; ++i) CHECK(seen.rets[i] == &arena[32 + i]);
    CHECK(seen.ws_bytes == (size_t)(1000 + nr - 1));      // the last result is the workspace
    next_code = JPCB_EINVAL;                               // a library error becomes an ffi::Error with the library's message
    CHECK(handlers[h](&f) != 0 && !f.result.success());
    CHECK(f.result.errc() == xla::ffi::ErrorCode::kInvalidArgument && f.result.message() == "recorded failure");
    next_code = JPCB_ECUDA;
    CHECK(handlers[h](&f) != 0 && f.result.errc() == xla::ffi::ErrorCode::kInternal);
    xla::ffi::MockCallFrame empty;                          // no results at all: refused before the library is called
    empty.stream = nullptr;
    const int calls = seen.calls;
    CHECK(handlers[h](&empty) != 0 && seen.calls == calls);
  }
  std::printf("shim ok: %d handlers\n", (int)JPCB_CALL_N_ENTRIES);
  return 0;
}
