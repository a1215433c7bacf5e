// MOCK of the small part of XLA's FFI C++ API that gpx_b200/csrc/ffi/ffi_shim.cc uses -- written for this repository
// (the image has no jaxlib, hence no real xla/ffi/api/ffi.h).  It exists only so that tests/test_abi_cpu.py can
// type-check the shim's handler bodies against include/gpxb200.h with `g++ -fsyntax-only`; it implements nothing
// and says nothing about the real library beyond the names and shapes the shim relies on:
//   ffi::Buffer<dtype>::typed_data() / dimensions(), ffi::ResultBuffer<dtype> (pointer-like), ffi::Error,
//   ffi::PlatformStream<T>, ffi::Ffi::Bind().Ctx<>().Arg<>().Attr<>(name).Ret<>(), XLA_FFI_DEFINE_HANDLER_SYMBOL.
#ifndef GPXB_MOCK_XLA_FFI_API_FFI_H_
#define GPXB_MOCK_XLA_FFI_API_FFI_H_

#include <cstddef>
#include <cstdint>
#include <string>

namespace xla {
namespace ffi {

enum DataType { F64, U32, S64 };

template <DataType dt> struct NativeType;
template <> struct NativeType<F64> { using type = double; };
template <> struct NativeType<U32> { using type = uint32_t; };
template <> struct NativeType<S64> { using type = int64_t; };

template <class T>
struct Span {
  const T* ptr = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return ptr[i]; }
};

template <DataType dt>
struct Buffer {
  using T = typename NativeType<dt>::type;
  T* data = nullptr;
  Span<int64_t> dims;
  T* typed_data() const { return data; }
  Span<int64_t> dimensions() const { return dims; }
};

template <class B>
struct Result {
  B buf;
  B* operator->() { return &buf; }
};
template <DataType dt>
using ResultBuffer = Result<Buffer<dt>>;

struct Error {
  static Error Success() { return Error(); }
  static Error Internal(const std::string&) { return Error(); }
};

template <class S>
struct PlatformStream {};

struct Binding {
  template <class T> Binding Ctx() const { return *this; }
  template <class T> Binding Arg() const { return *this; }
  template <class T> Binding Attr(const char*) const { return *this; }
  template <class T> Binding Ret() const { return *this; }
};
struct Ffi {
  static Binding Bind() { return Binding(); }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding) \
  extern "C" const void* symbol() {                          \
    (void)(binding);                                         \
    return reinterpret_cast<const void*>(&impl);             \
  }

#endif  // GPXB_MOCK_XLA_FFI_API_FFI_H_
