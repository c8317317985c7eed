// TEST STUB -- not jaxlib's header.  A minimal stand-in for xla/ffi/api/ffi.h with the names csrc/mtx_jax_ffi.cc uses, so that the
// compile-guarded XLA FFI handlers can be type-checked in an image without jaxlib (tests/test_abi.py compiles the file against
// it with -fsyntax-only).  XLA_FFI_DEFINE_HANDLER_SYMBOL static_asserts that the handler is invocable with the bound context,
// arguments, results and attributes in order, returning xla::ffi::Error.  Follows the public API of jaxlib 0.6 (xla::ffi::Buffer,
// ResultBuffer, Ffi::Bind().Ctx/Arg/Ret/Attr, PlatformStream); nothing here is used at run time.
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>

namespace xla {
namespace ffi {

enum DataType { BF16, F32, S32, U32, U8 };
enum class ErrorCode { kInternal, kInvalidArgument };

class Error {
 public:
  Error(ErrorCode, const char*) {}
  static Error Success() { return Error(ErrorCode::kInternal, ""); }
};

template <DataType T> struct NativeType { using type = uint16_t; };
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<U32> { using type = uint32_t; };
template <> struct NativeType<U8> { using type = uint8_t; };

struct Dims {
  const int64_t* p = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  int64_t operator[](size_t i) const { return p[i]; }
};

template <DataType T>
class Buffer {
 public:
  Dims dimensions() const { return {}; }
  void* untyped_data() const { return nullptr; }
  typename NativeType<T>::type* typed_data() const { return nullptr; }
  size_t element_count() const { return 0; }
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
 private:
  T value_;
};
template <DataType T> using ResultBuffer = Result<Buffer<T>>;

template <typename T> struct PlatformStream {};

template <typename... Ts>
struct Binding {
  template <typename T> struct CtxType;
  template <typename S> struct CtxType<PlatformStream<S>> { using type = S; };
  template <typename T> Binding<Ts..., typename CtxType<T>::type> Ctx() const { return {}; }
  template <typename T> Binding<Ts..., T> Arg() const { return {}; }
  template <typename T> Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T> Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename F> static constexpr bool Matches() { return std::is_invocable_r<Error, F, Ts...>::value; }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                                                            \
  static_assert(decltype(binding)::Matches<decltype(&impl)>(), #impl " does not match the bound operands of " #symbol); \
  extern "C" {                                                                                                          \
  const void* symbol = reinterpret_cast<const void*>(&impl);                                                            \
  }                                                                                                                     \
  static_assert(true, "")
