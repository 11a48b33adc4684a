// Shared plumbing of the extern "C" translation units.
#pragma once
#include <string>

#include "dla.h"

struct t4b_ctx {
    t4b::dla::Ctx* c;
};

std::string& t4b_last_error_ref();

#define T4B_TRY try {
#define T4B_CATCH                                                      \
    }                                                                  \
    catch (const t4b::Error& e) {                                      \
        t4b_last_error_ref() = e.what();                               \
        return (int)e.code;                                            \
    }                                                                  \
    catch (const std::exception& e) {                                  \
        t4b_last_error_ref() = e.what();                               \
        return T4B_INTERNAL;                                           \
    }                                                                  \
    catch (...) {                                                      \
        t4b_last_error_ref() = "unknown error";                        \
        return T4B_INTERNAL;                                           \
    }                                                                  \
    return T4B_OK;

inline void require_ctx(t4b_ctx* ctx) {
    if (!ctx || !ctx->c) throw t4b::Error(t4b::ST_INVALID_ARGUMENT, "null context");
    // contexts are per device: make that device current for everything this call launches or allocates
    t4b::dla::make_current(ctx->c);
}
inline t4b::DType to_dtype(int d) {
    if (d == 0) return t4b::F64;
    if (d == 1) return t4b::C64;
    throw t4b::Error(t4b::ST_INVALID_ARGUMENT, "dtype must be T4B_F64 or T4B_C64");
}
