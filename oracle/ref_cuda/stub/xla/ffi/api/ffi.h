// Stand-in for jaxlib's xla/ffi/api/ffi.h -- TEST INFRASTRUCTURE ONLY.
//
// Two reference sources (brainevent/_csr/binary_csrmv_hybrid.cu:67 and
// _csr/binary_indexed_csrmv_hybrid.cu) include the XLA FFI header only to use
// xla::ffi::Error as an internal return type and, through
// include/brainevent/dtypes.h:23-45, the xla::ffi::DataType tags.  jaxlib is not in this
// image, so this header declares just those two names; nothing here is reference code.
#pragma once
#include <string>
#include <utility>
namespace xla { namespace ffi {
enum class DataType : int {
    INVALID = 0, PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16, C64, C128
};
class Error {
public:
    Error() = default;
    static Error Success() { return Error(); }
    static Error Internal(std::string m) { Error e; e.ok_ = false; e.msg_ = std::move(m); return e; }
    bool success() const { return ok_; }
    bool failure() const { return !ok_; }
    const std::string& message() const { return msg_; }
private:
    bool ok_ = true;
    std::string msg_;
};
}}  // namespace xla::ffi
