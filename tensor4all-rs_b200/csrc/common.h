// t4b — B200-native dense tensor-train engine: shared host-side definitions.
//
// Everything above this header is plain C++17 (no CUDA types leak out of
// csrc/kernels/): the sweep drivers in csrc/host/ talk to the device only
// through the launchers declared in kernels.h.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace t4b {

enum DType : int { F64 = 0, C64 = 1 };
inline size_t dtype_size(DType d) { return d == F64 ? 8 : 16; }

struct cplx {
    double re, im;
};

// Status codes mirror the reference C-API conventions
// (docs/CAPI_DESIGN.md:24-107: status enum + thread-local last-error string).
enum Status : int {
    ST_OK = 0,
    ST_INVALID_ARGUMENT = 1,
    ST_CUDA_ERROR = 2,
    ST_NOT_CONVERGED = 3,
    ST_UNSUPPORTED = 4,
    ST_INTERNAL = 5,
};

struct Error : std::runtime_error {
    Status code;
    Error(Status c, const std::string& msg) : std::runtime_error(msg), code(c) {}
};

#define T4B_REQUIRE(cond, msg)                                                              \
    do {                                                                                    \
        if (!(cond)) throw ::t4b::Error(::t4b::ST_INVALID_ARGUMENT, std::string(msg));     \
    } while (0)

// A composite index: up to kMaxGroupDims tensor axes fused into one matrix
// index, first axis fastest.  offset(i) = sum_j digit_j(i) * stride_j.
constexpr int kMaxGroupDims = 6;
struct Group {
    int nd = 0;
    int64_t dim[kMaxGroupDims] = {1, 1, 1, 1, 1, 1};
    int64_t str[kMaxGroupDims] = {0, 0, 0, 0, 0, 0};
    int64_t size() const {
        int64_t s = 1;
        for (int i = 0; i < nd; ++i) s *= dim[i];
        return s;
    }
};

}  // namespace t4b
