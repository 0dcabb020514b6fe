// TEST INFRASTRUCTURE -- a stand-in for jaxlib's xla/ffi/api/ffi.h, NOT the real header and not part of the product.
//
// jaxlib is not installable in this image, so jpc_b200/csrc/xla_ffi_shim.cc can never be compiled against the real XLA FFI
// API here.  This file restates, from the published API, exactly the surface the shim touches -- ffi::Span, ffi::AnyBuffer,
// ffi::RemainingArgs / RemainingRets with get<T>(i) returning ErrorOr<T> / ErrorOr<Result<T>>, ffi::Error / ErrorCode,
// ffi::PlatformStream<T>, the Ffi::Bind() builder chain and XLA_FFI_DEFINE_HANDLER_SYMBOL -- with trivial bodies, so that
// tests/test_xla_shim_mock.py can (1) prove the shim is well-formed C++ against that API shape and (2) drive every handler
// with fake buffers and check what reaches jpcb_call.  A handler symbol here takes a MockCallFrame* instead of XLA's
// XLA_FFI_CallFrame*.  If the real header disagrees with this restatement the shim needs the same fix there; this mock only
// removes the "never seen by a compiler" risk.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace xla { namespace ffi {

enum class ErrorCode { kOk = 0, kInvalidArgument = 3, kInternal = 13 };

class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
  ErrorCode errc() const { return code_; }
  const std::string& message() const { return msg_; }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <typename T>
class Span {
 public:
  Span(T* p, size_t n) : p_(p), n_(n) {}
  T* begin() const { return p_; }
  T* end() const { return p_ + n_; }
  size_t size() const { return n_; }
  T& operator[](size_t i) const { return p_[i]; }
 private:
  T* p_; size_t n_;
};

class AnyBuffer {
 public:
  AnyBuffer(void* p, size_t bytes) : p_(p), bytes_(bytes) {}
  void* untyped_data() const { return p_; }
  size_t size_bytes() const { return bytes_; }
 private:
  void* p_; size_t bytes_;
};

template <typename T>
class Result {
 public:
  explicit Result(T v) : v_(v) {}
  T& operator*() { return v_; }
  T* operator->() { return &v_; }
 private:
  T v_;
};

template <typename T>
class ErrorOr {
 public:
  ErrorOr(T v) : ok_(true), v_(std::move(v)) {}
  bool has_value() const { return ok_; }
  T& operator*() { return v_; }
  T* operator->() { return &v_; }
 private:
  bool ok_; T v_;
};

class RemainingArgs {
 public:
  explicit RemainingArgs(const std::vector<AnyBuffer>* b) : b_(b) {}
  size_t size() const { return b_->size(); }
  template <typename T> ErrorOr<T> get(size_t i) const { return ErrorOr<T>((*b_)[i]); }
 private:
  const std::vector<AnyBuffer>* b_;
};

class RemainingRets {
 public:
  explicit RemainingRets(const std::vector<AnyBuffer>* b) : b_(b) {}
  size_t size() const { return b_->size(); }
  template <typename T> ErrorOr<Result<T>> get(size_t i) const { return ErrorOr<Result<T>>(Result<T>((*b_)[i])); }
 private:
  const std::vector<AnyBuffer>* b_;
};

template <typename T> struct PlatformStream {};

// what the mock hands a handler: XLA's call frame reduced to the four things the shim binds
struct MockCallFrame {
  void* stream;
  std::vector<uint8_t> desc;
  std::vector<AnyBuffer> args, rets;
  Error result;
};

// Ffi::Bind().Ctx<PlatformStream<S>>().Attr<Span<const uint8_t>>("desc").RemainingArgs().RemainingRets(): the mock keeps
// only the stream type S, so the macro below can call impl(S, Span, RemainingArgs, RemainingRets).
template <typename S>
struct MockBinding {
  template <typename A> MockBinding Attr(const char*) const { return *this; }
  MockBinding RemainingArgs() const { return *this; }
  MockBinding RemainingRets() const { return *this; }
  template <typename Fn> Error Call(Fn fn, MockCallFrame* f) const {
    return fn(reinterpret_cast<S>(f->stream), Span<const uint8_t>(f->desc.data(), f->desc.size()), ffi::RemainingArgs(&f->args),
              ffi::RemainingRets(&f->rets));
  }
};
struct MockBindStart {
  template <typename C> auto Ctx() const { return ctx(static_cast<C*>(nullptr)); }
  template <typename S> static MockBinding<S> ctx(PlatformStream<S>*) { return {}; }
};
struct Ffi { static MockBindStart Bind() { return {}; } };

}}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                  \
  extern "C" int symbol(::xla::ffi::MockCallFrame* frame) {                  \
    frame->result = (binding).Call(impl, frame);                             \
    return frame->result.success() ? 0 : (int)frame->result.errc();          \
  }
