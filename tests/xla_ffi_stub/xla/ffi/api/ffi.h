// COMPILE-CHECK STUB, test infrastructure only -- NOT the XLA header.
//
// jaxlib (and with it xla/ffi/api/ffi.h) is not installable in this project's images, so cmcd_b200/csrc/xla_ffi.cc is never built
// against the real thing here.  This stub declares the small part of the public XLA-FFI C++ API surface that file uses --
// Buffer / ResultBuffer accessors, Error, the Ffi::Bind() builder and XLA_FFI_DEFINE_HANDLER_SYMBOL -- with just enough type
// machinery that `g++ -fsyntax-only` catches signature drift: the binding's Ctx / Arg / Ret / Attr list must be callable on the
// implementation function, argument for argument (tests/test_capi_symbols.py::test_xla_ffi_shim_compiles_against_stub).
// It generates no code and is never linked into the product.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>

namespace xla {
namespace ffi {

enum DataType { F32, S32, U8 };
enum class ErrorCode { kInvalidArgument, kInternal };

class Error {
   public:
    Error() = default;
    Error(ErrorCode, std::string) {}
    static Error Success() { return Error(); }
};

template <DataType dt> struct NativeType;
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<U8> { using type = uint8_t; };

struct Dims {
    const int64_t* p = nullptr;
    int64_t operator[](size_t i) const { return p[i]; }
    size_t size() const { return 0; }
};

template <DataType dt>
class Buffer {
   public:
    using T = typename NativeType<dt>::type;
    T* typed_data() const { return nullptr; }
    size_t element_count() const { return 0; }
    Dims dimensions() const { return Dims(); }
};

template <typename T>
class Result {
   public:
    T* operator->() { return &v_; }
   private:
    T v_;
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

template <typename T> struct PlatformStream {};
template <typename C> struct CtxType;
template <typename T> struct CtxType<PlatformStream<T>> { using type = T; };

template <typename... Ts>
struct Binding {
    template <typename C> Binding<Ts..., typename CtxType<C>::type> Ctx() { return {}; }
    template <typename A> Binding<Ts..., A> Arg() { return {}; }
    template <typename R> Binding<Ts..., Result<R>> Ret() { return {}; }
    template <typename A> Binding<Ts..., A> Attr(const char*) { return {}; }
    template <typename F> static constexpr bool Matches = std::is_invocable_r_v<Error, F, Ts...>;
};

struct Ffi {
    static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, ...)                                                                  \
    static_assert(std::remove_reference_t<decltype(__VA_ARGS__)>::template Matches<decltype(&impl)>,                    \
                  #name ": the FFI binding's Ctx/Arg/Ret/Attr list does not match the implementation's parameter list"); \
    extern "C" void* name(void*) { return nullptr; }
