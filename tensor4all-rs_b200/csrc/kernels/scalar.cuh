// Real / Complex64 scalar helpers and DMMA fragment helpers shared by the factorisation kernels.
#pragma once
#include "ctx.cuh"

namespace t4b {
namespace dla {

template <bool CPLX> struct Sc;
template <> struct Sc<false> {
    typedef double T;
    __device__ __forceinline__ static T zero() { return 0.0; }
    __device__ __forceinline__ static T one() { return 1.0; }
    __device__ __forceinline__ static T conj(T a) { return a; }
    __device__ __forceinline__ static T mul(T a, T b) { return a * b; }
    __device__ __forceinline__ static T add(T a, T b) { return a + b; }
    __device__ __forceinline__ static T sub(T a, T b) { return a - b; }
    __device__ __forceinline__ static T neg(T a) { return -a; }
    __device__ __forceinline__ static T scale(T a, double s) { return a * s; }
    __device__ __forceinline__ static double abs2(T a) { return a * a; }
    __device__ __forceinline__ static double real(T a) { return a; }
    __device__ __forceinline__ static T from_real(double r) { return r; }
    __device__ __forceinline__ static T shfl_xor(T v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
};
template <> struct Sc<true> {
    typedef double2 T;
    __device__ __forceinline__ static T zero() { return make_double2(0.0, 0.0); }
    __device__ __forceinline__ static T one() { return make_double2(1.0, 0.0); }
    __device__ __forceinline__ static T conj(T a) { return make_double2(a.x, -a.y); }
    __device__ __forceinline__ static T mul(T a, T b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
    __device__ __forceinline__ static T add(T a, T b) { return make_double2(a.x + b.x, a.y + b.y); }
    __device__ __forceinline__ static T sub(T a, T b) { return make_double2(a.x - b.x, a.y - b.y); }
    __device__ __forceinline__ static T neg(T a) { return make_double2(-a.x, -a.y); }
    __device__ __forceinline__ static T scale(T a, double s) { return make_double2(a.x * s, a.y * s); }
    __device__ __forceinline__ static double abs2(T a) { return a.x * a.x + a.y * a.y; }
    __device__ __forceinline__ static double real(T a) { return a.x; }
    __device__ __forceinline__ static T from_real(double r) { return make_double2(r, 0.0); }
    __device__ __forceinline__ static T shfl_xor(T v, int o) {
        return make_double2(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
    }
};

template <bool CPLX>
__device__ __forceinline__ typename Sc<CPLX>::T warp_sum_t(typename Sc<CPLX>::T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = Sc<CPLX>::add(v, Sc<CPLX>::shfl_xor(v, o));
    return v;
}

template <bool CPLX>
__device__ __forceinline__ typename Sc<CPLX>::T shfl_t(typename Sc<CPLX>::T v, int src) {
    if constexpr (CPLX) {
        return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
    } else {
        return __shfl_sync(0xffffffffu, v, src);
    }
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// acc(8x8 fragment: this thread holds (row grp, cols 2*tig, 2*tig+1)) += op(a) * b
// a: A fragment element (row grp, k tig); b: B fragment element (k tig, col grp).
template <bool CPLX, bool CONJA>
__device__ __forceinline__ void mma_frag(typename Sc<CPLX>::T (&acc)[2], typename Sc<CPLX>::T a,
                                         typename Sc<CPLX>::T b) {
    if constexpr (CPLX) {
        const double ar = a.x, ai = CONJA ? -a.y : a.y;
        dmma884(acc[0].x, acc[1].x, ar, b.x);
        dmma884(acc[0].x, acc[1].x, -ai, b.y);
        dmma884(acc[0].y, acc[1].y, ar, b.y);
        dmma884(acc[0].y, acc[1].y, ai, b.x);
    } else {
        dmma884(acc[0], acc[1], a, b);
    }
}

}  // namespace dla
}  // namespace t4b
