// Minimal stand-in for jaxlib's "xla/ffi/api/ffi.h" -- TEST INFRASTRUCTURE ONLY.
//
// jax / jaxlib are not installed in this image, so the real header is absent.
// This stub declares just the subset of the XLA FFI C++ API that
// liesel_ptm_b200/csrc/ptm_xla_ffi.cc uses, with the same names and shapes, so that the TU
// can be type-checked here (tests/test_host.py: g++ -fsyntax-only -DPTM_FFI_STUB ...).
// Binding<...>::To() static_asserts that the handler is invocable with the C++ types the
// binding describes -- the check the real header performs.  It cannot run anything.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla::ffi {

enum DataType { U8, S32, S64, F32, F64 };
template <DataType> struct NativeTypeOf;
template <> struct NativeTypeOf<U8> { using type = uint8_t; };
template <> struct NativeTypeOf<S32> { using type = int32_t; };
template <> struct NativeTypeOf<S64> { using type = int64_t; };
template <> struct NativeTypeOf<F32> { using type = float; };
template <> struct NativeTypeOf<F64> { using type = double; };

template <typename T>
struct Span {
  const T* ptr = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return ptr[i]; }
};

template <DataType DT>
class Buffer {
 public:
  using T = typename NativeTypeOf<DT>::type;
  T* typed_data() const { return data_; }
  Span<int64_t> dimensions() const { return dims_; }
  size_t element_count() const { return count_; }

 private:
  T* data_ = nullptr;
  Span<int64_t> dims_;
  size_t count_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};
template <DataType DT> using ResultBuffer = Result<Buffer<DT>>;

enum class ErrorCode { kOk, kInvalidArgument, kInternal };
class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <typename T> struct PlatformStream {};

namespace internal {
template <typename T> struct CtxType;
template <typename T> struct CtxType<PlatformStream<T>> { using type = T; };
template <typename T> struct RetType { using type = Result<T>; };
struct Handler { XLA_FFI_Error* Call(XLA_FFI_CallFrame*) { return nullptr; } };
}  // namespace internal

template <typename... Ts>
struct Binding {
  template <typename T> Binding<Ts..., typename internal::CtxType<T>::type> Ctx() { return {}; }
  template <typename T> Binding<Ts..., T> Arg() { return {}; }
  template <typename T> Binding<Ts..., T> Attr(const char*) { return {}; }
  template <typename T> Binding<Ts..., typename internal::RetType<T>::type> Ret() { return {}; }
  template <typename F>
  internal::Handler* To(F) {
    static_assert(std::is_invocable_r_v<Error, F, Ts...>,
                  "handler signature does not match the FFI binding");
    static internal::Handler h;
    return &h;
  }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(fn, impl, binding)               \
  extern "C" XLA_FFI_Error* fn(XLA_FFI_CallFrame* call_frame) {        \
    static auto* handler = (binding).To(impl);                         \
    return handler->Call(call_frame);                                  \
  }
