// jacobi_eig.cuh - Hermitian Jacobi eigen-solve of the 32 x 32 pair Gram block in shared memory (the serial part of
// every item of jacobi_persistent_kernel, svd.cu).  In its own header so that tools/probe_eig.cu can time it alone.
#pragma once
#include "scalar.cuh"

namespace t4b {
namespace dla {
namespace {

constexpr int JT = 256;    // threads per CTA
constexpr int WP = 37;     // pitch of the W matrix in shared memory: odd, so that the column-pair updates of the inner eigen-solve (16 lanes = 16 columns, same row) are free of bank conflicts (36 made them 2x conflicted; the 32 one-off fragment loads of the update pass pay 1.5x instead)
constexpr int GP = 33;     // pitch of the G matrix in shared memory

// Hermitian Jacobi on the 32 x 32 Gram block held in shared memory (Gs, ping-pong copy Gs2), rotations
// accumulated into Ws (initialised to the identity by the caller).  16 disjoint rotations per step;
// every thread owns one 2 x 2 block of G' = Ja^H G Jb.  full_inner: 31-step round-robin over all 32
// indices; otherwise the 16-step bipartite ordering (cross pairs block I x block J only).
// Called by all JT threads; ends with a __syncthreads().
template <bool CPLX>
__device__ __forceinline__ typename Sc<CPLX>::T* jacobi_eig32(typename Sc<CPLX>::T* Gs, typename Sc<CPLX>::T* Gs2,
                                             typename Sc<CPLX>::T* Ws, typename Sc<CPLX>::T* rot_ph,
                                             double* rot_c, double* rot_s, int full_inner, double tol_rot,
                                             int tid, int inner = 1, double zthr = 0.0) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    T* Gcur = Gs;
    T* Gnxt = Gs2;
    const int nrr = full_inner ? 31 : 16 * inner;
    const int ta = tid >> 4, tb = tid & 15;   // row pair / column pair owned by this thread
    const double tol2 = tol_rot * tol_rot;
    for (int rr = 0; rr < nrr; ++rr) {
        int pa, qa, pb, qb;
        if (full_inner) {
            if (ta == 0) { pa = rr; qa = 31; } else { pa = (rr + ta) % 31; qa = (rr - ta + 62) % 31; }
            if (pa > qa) { int t = pa; pa = qa; qa = t; }
            if (tb == 0) { pb = rr; qb = 31; } else { pb = (rr + tb) % 31; qb = (rr - tb + 62) % 31; }
            if (pb > qb) { int t = pb; pb = qb; qb = t; }
        } else {
            pa = ta; qa = 16 + ((ta + rr) & 15);
            pb = tb; qb = 16 + ((tb + rr) & 15);
        }
        // 16 threads compute the 16 rotations of this step (pair t = tid):
        // J = [[c, s], [-s e^{-i phi}, c e^{-i phi}]], J^H [[aa,g],[g*,bb]] J diagonal
        if (tid < 16) {
            int pp, qq;
            if (full_inner) {
                if (tid == 0) { pp = rr; qq = 31; } else { pp = (rr + tid) % 31; qq = (rr - tid + 62) % 31; }
                if (pp > qq) { int t = pp; pp = qq; qq = t; }
            } else {
                pp = tid; qq = 16 + ((tid + rr) & 15);
            }
            double c = 1.0, sn = 0.0;
            T ph = S::one();
            const double aa = S::real(Gcur[pp * GP + pp]), bb = S::real(Gcur[qq * GP + qq]);
            const T g = Gcur[pp * GP + qq];
            const double g2 = S::abs2(g);
            if (g2 > 0.0 && aa > zthr && bb > zthr && g2 > tol2 * aa * bb) {
                const double d = 0.5 * (bb - aa);
                if constexpr (CPLX) {
                    const double inv_absg = rsqrt(g2);
                    const double absg = g2 * inv_absg;
                    const double x = d * d + g2;
                    const double h = x * rsqrt(x);
                    const double t = (d >= 0.0 ? absg : -absg) / (fabs(d) + h);
                    c = rsqrt(1.0 + t * t);
                    sn = c * t;
                    ph = S::scale(S::conj(g), inv_absg);   // e^{-i phi}
                } else {
                    // t = sign(d) |g| / (|d| + sqrt(d^2 + g^2)) only steers the convergence, so it is
                    // evaluated in fp32 on exponent-normalised operands; c = (1 + t^2)^(-1/2) must make
                    // the rotation orthogonal to fp64 accuracy: fp32 seed + two Newton steps.
                    const double ag = fabs(g), ad = fabs(d);
                    const double mx = ag > ad ? ag : ad;
                    const int ex = (__double2hiint(mx) >> 20) & 0x7ff;
                    const double sc = __hiloint2double((2046 - ex) << 20, 0);     // mx * sc in [1, 2)
                    const float fd = (float)(ad * sc), fg = (float)(ag * sc);
                    const float fh = sqrtf(fd * fd + fg * fg);
                    const float ft = __fdividef(fg, fd + fh);
                    const double t = d >= 0.0 ? (double)ft : -(double)ft;
                    const double x = 1.0 + t * t;
                    double y = (double)rsqrtf((float)x);
                    y = y * (1.5 - 0.5 * x * y * y);
                    y = y * (1.5 - 0.5 * x * y * y);
                    c = y;
                    sn = c * t;
                    ph = g >= 0.0 ? 1.0 : -1.0;
                }
            }
            rot_c[tid] = c; rot_s[tid] = sn; rot_ph[tid] = ph;
        }
        __syncthreads();
        const double ca = rot_c[ta], sa = rot_s[ta], cb = rot_c[tb], sb = rot_s[tb];
        const T pha = rot_ph[ta], phb = rot_ph[tb];
        // G' = Ja^H G Jb on the 2 x 2 block owned by this thread
        const T g00 = Gcur[pa * GP + pb], g01 = Gcur[pa * GP + qb];
        const T g10 = Gcur[qa * GP + pb], g11 = Gcur[qa * GP + qb];
        const T cpa = S::conj(pha);
        const T e10 = S::mul(cpa, g10), e11 = S::mul(cpa, g11);
        const T r00 = S::sub(S::scale(g00, ca), S::scale(e10, sa));
        const T r01 = S::sub(S::scale(g01, ca), S::scale(e11, sa));
        const T r10 = S::add(S::scale(g00, sa), S::scale(e10, ca));
        const T r11 = S::add(S::scale(g01, sa), S::scale(e11, ca));
        const T f01 = S::mul(r01, phb), f11 = S::mul(r11, phb);
        Gnxt[pa * GP + pb] = S::sub(S::scale(r00, cb), S::scale(f01, sb));
        Gnxt[pa * GP + qb] = S::add(S::scale(r00, sb), S::scale(f01, cb));
        Gnxt[qa * GP + pb] = S::sub(S::scale(r10, cb), S::scale(f11, sb));
        Gnxt[qa * GP + qb] = S::add(S::scale(r10, sb), S::scale(f11, cb));
        // W <- W Jb for the two rows owned by this thread (exclusive ownership: in place)
#pragma unroll
        for (int rrow = 0; rrow < 2; ++rrow) {
            const int i = ta * 2 + rrow;
            const T wp = Ws[pb * WP + i], wq = Ws[qb * WP + i];
            const T fq = S::mul(wq, phb);
            Ws[pb * WP + i] = S::sub(S::scale(wp, cb), S::scale(fq, sb));
            Ws[qb * WP + i] = S::add(S::scale(wp, sb), S::scale(fq, cb));
        }
        __syncthreads();
        T* tmp = Gcur; Gcur = Gnxt; Gnxt = tmp;
    }
    return Gcur;
}


// Rotation of the pivot (aa, g; conj(g), bb): J = [[c, s], [-s e^{-i phi}, c e^{-i phi}]] with J^H (..) J diagonal.
// Returns tg = t |g| (t = tan of the rotation angle): the rotated diagonal is (aa - tg, bb + tg).
template <bool CPLX>
__device__ __forceinline__ void jacobi_rotation(double aa, double bb, typename Sc<CPLX>::T g, double tol2,
                                                double& c, double& sn, typename Sc<CPLX>::T& ph, double& tg,
                                                double zthr = 0.0) {
    typedef Sc<CPLX> S;
    c = 1.0; sn = 0.0; ph = S::one(); tg = 0.0;
    const double g2 = S::abs2(g);
    if (g2 > 0.0 && aa > zthr && bb > zthr && g2 > tol2 * aa * bb) {
        const double d = 0.5 * (bb - aa);
        if constexpr (CPLX) {
            const double inv_absg = rsqrt(g2);
            const double absg = g2 * inv_absg;
            const double x = d * d + g2;
            const double h = x * rsqrt(x);
            const double t = (d >= 0.0 ? absg : -absg) / (fabs(d) + h);
            c = rsqrt(1.0 + t * t);
            sn = c * t;
            ph = S::scale(S::conj(g), inv_absg);
            tg = t * absg;
        } else {
            // (c, s) = (u, |g|) / sqrt(u^2 + g^2), u = |d| + sqrt(d^2 + g^2), evaluated in fp32 on exponent-
            // normalised operands (two MUFU ops, no division): the angle only steers the convergence.  The
            // pair is then renormalised in fp64, (c, s) *= 1 - e/2 + 3 e^2 / 8 with e = c^2 + s^2 - 1 ~ 1e-7,
            // which makes the rotation orthogonal to ~e^3.
            const double ag = fabs(g), ad = fabs(d);
            const double mx = ag > ad ? ag : ad;
            const int ex = (__double2hiint(mx) >> 20) & 0x7ff;
            const double sc = __hiloint2double((2046 - ex) << 20, 0);     // mx * sc in [1, 2)
            const float fd = (float)(ad * sc), fg = (float)(ag * sc);
            const float fu = fd + sqrtf(fd * fd + fg * fg);
            const float rho = rsqrtf(fu * fu + fg * fg);
            double cd = (double)(fu * rho), sd = (double)(fg * rho);
            if (fg == 0.0f) { cd = 1.0; sd = ag / (2.0 * ad); }          // |g| / |d| below fp32 range: tiny angle
            const double e = fma(cd, cd, fma(sd, sd, -1.0));
            const double corr = fma(e, fma(e, 0.375, -0.5), 1.0);
            cd *= corr; sd *= corr;
            c = cd;
            sn = d >= 0.0 ? sd : -sd;
            ph = g >= 0.0 ? 1.0 : -1.0;
            tg = fma(sn * sn, aa - bb, 2.0 * c * sn * ag);               // aa - (c^2 aa - 2 c s |g| + s^2 bb)
        }
    }
}

// Bipartite (cross pairs only) inner sweeps with the rotation warp running ONE STEP AHEAD of the apply warps:
// while warps 1..7 apply the rotations of step r (G' = J^H G J, W <- W J), warp 0 already derives the pivots of
// step r+1 - G'[p][p] = aa - tg, G'[q'][q'] = bb' + tg' of the neighbouring pair, G'[p][q'] from four entries of
// G and the two rotations involved - and computes the next rotations.  One __syncthreads per step.
template <bool CPLX>
__device__ __forceinline__ typename Sc<CPLX>::T* jacobi_eig32_pipelined(
    typename Sc<CPLX>::T* Gs, typename Sc<CPLX>::T* Gs2, typename Sc<CPLX>::T* Ws, typename Sc<CPLX>::T* rot_ph,
    double* rot_c, double* rot_s, double tol_rot, int tid, int inner, double zthr) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    T* Gcur = Gs;
    T* Gnxt = Gs2;
    const int nrr = 16 * inner;
    const double tol2 = tol_rot * tol_rot;
    const int lane = tid & 31, warp = tid >> 5;
    // rotation state of pair t = lane (warp 0, lanes 0..15)
    double c = 1.0, sn = 0.0, tg = 0.0, aa = 0.0, bb = 0.0;
    T ph = S::one();
    if (warp == 0 && lane < 16) {
        const int pp = lane, qq = 16 + lane;
        aa = S::real(Gcur[pp * GP + pp]); bb = S::real(Gcur[qq * GP + qq]);
        jacobi_rotation<CPLX>(aa, bb, Gcur[pp * GP + qq], tol2, c, sn, ph, tg, zthr);
        rot_c[lane] = c; rot_s[lane] = sn; rot_ph[lane] = ph;
    }
    __syncthreads();
    for (int rr = 0; rr < nrr; ++rr) {
        const int cur = (rr & 1) * 16, nxt = 16 - cur;
        if (warp == 0) {
            if (lane < 16 && rr + 1 < nrr) {
                // pivots of step rr+1 for pair (p, q'), q' = partner of lane+1 at step rr
                const int nb = (lane + 1) & 15;
                const int pp = lane, qq = 16 + ((lane + rr) & 15);
                const int pn = nb, qn = 16 + ((nb + rr) & 15);
                const T gpp = Gcur[pp * GP + pn], gpq = Gcur[pp * GP + qn];
                const T gqp = Gcur[qq * GP + pn], gqq = Gcur[qq * GP + qn];
                const double cn = __shfl_sync(0x0000ffffu, c, nb), snn = __shfl_sync(0x0000ffffu, sn, nb);
                const double bbn = __shfl_sync(0x0000ffffu, bb + tg, nb);
                T phn;
                if constexpr (CPLX) phn = make_double2(__shfl_sync(0x0000ffffu, ph.x, nb), __shfl_sync(0x0000ffffu, ph.y, nb));
                else phn = __shfl_sync(0x0000ffffu, ph, nb);
                // column combination with the neighbour's rotation, then row combination with the own one
                const T yp = S::add(S::scale(gpp, snn), S::scale(S::mul(gpq, phn), cn));
                const T yq = S::add(S::scale(gqp, snn), S::scale(S::mul(gqq, phn), cn));
                const T gnew = S::sub(S::scale(yp, c), S::scale(S::mul(S::conj(ph), yq), sn));
                aa = aa - tg; bb = bbn;
                jacobi_rotation<CPLX>(aa, bb, gnew, tol2, c, sn, ph, tg, zthr);
                rot_c[nxt + lane] = c; rot_s[nxt + lane] = sn; rot_ph[nxt + lane] = ph;
            }
        } else {
            // 256 2 x 2 blocks over the 224 threads of warps 1..7
            for (int blk = tid - 32; blk < 256; blk += 224) {
                const int ta = blk >> 4, tb = blk & 15;
                const int pa = ta, qa = 16 + ((ta + rr) & 15);
                const int pb = tb, qb = 16 + ((tb + rr) & 15);
                const double ca = rot_c[cur + ta], sa = rot_s[cur + ta], cb = rot_c[cur + tb], sb = rot_s[cur + tb];
                const T pha = rot_ph[cur + ta], phb = rot_ph[cur + tb];
                const T g00 = Gcur[pa * GP + pb], g01 = Gcur[pa * GP + qb];
                const T g10 = Gcur[qa * GP + pb], g11 = Gcur[qa * GP + qb];
                const T cpa = S::conj(pha);
                const T e10 = S::mul(cpa, g10), e11 = S::mul(cpa, g11);
                const T r00 = S::sub(S::scale(g00, ca), S::scale(e10, sa));
                const T r01 = S::sub(S::scale(g01, ca), S::scale(e11, sa));
                const T r10 = S::add(S::scale(g00, sa), S::scale(e10, ca));
                const T r11 = S::add(S::scale(g01, sa), S::scale(e11, ca));
                const T f01 = S::mul(r01, phb), f11 = S::mul(r11, phb);
                Gnxt[pa * GP + pb] = S::sub(S::scale(r00, cb), S::scale(f01, sb));
                Gnxt[pa * GP + qb] = S::add(S::scale(r00, sb), S::scale(f01, cb));
                Gnxt[qa * GP + pb] = S::sub(S::scale(r10, cb), S::scale(f11, sb));
                Gnxt[qa * GP + qb] = S::add(S::scale(r10, sb), S::scale(f11, cb));
#pragma unroll
                for (int rrow = 0; rrow < 2; ++rrow) {
                    const int i = ta * 2 + rrow;
                    const T wp = Ws[pb * WP + i], wq = Ws[qb * WP + i];
                    const T fq = S::mul(wq, phb);
                    Ws[pb * WP + i] = S::sub(S::scale(wp, cb), S::scale(fq, sb));
                    Ws[qb * WP + i] = S::add(S::scale(wp, sb), S::scale(fq, cb));
                }
            }
        }
        __syncthreads();
        T* tmp = Gcur; Gcur = Gnxt; Gnxt = tmp;
    }
    return Gcur;
}
// ---- v2 (real case) ---------------------------------------------------------------------------------------------------
// Measured with tools/probe_eig.cu (B200): the one-step-ahead form above costs 865 cycles per step - a 680-cycle serial
// chain in the rotation warp next to apply warps that are bound by shared-memory bandwidth (G and W both live there) and
// by two serialised blocks per thread.  v2:
//   * plain rotations J = [[c, s], [-s, c]] with a SIGNED s: no phase factor anywhere;
//   * the rotation warp publishes the fp32-accurate SEEDS (cd, sd) and steers with them; the fp64 renormalisation
//     (c, s) = (cd, sd) (1 - e/2 + 3 e^2/8), e = cd^2 + sd^2 - 1, is evaluated by the consumers (identical arithmetic
//     on identical inputs, so G and W see the same rotation) - five dependent fp64 operations leave the serial chain;
//   * G' = J^T G J only on the upper block triangle of the symmetric G with mirrored stores: 136 two-by-two blocks, ONE
//     per thread of warps 1..5;
//   * W lives in the registers of warp 7 (one row per lane, the q-half rotated by one register per step so that every
//     index is a compile-time constant) and is written to shared memory once, in the last step.
// Rotation seeds of the pivot (aa, g; g, bb): cd, sd (|cd^2 + sd^2 - 1| ~ 1e-7, sd signed) and tg ~ t g (the rotated
// diagonal is (aa - tg, bb + tg); steering only).
__device__ __forceinline__ void jacobi_rotation_seed(double aa, double bb, double g, double tol2, double& cd, double& sd,
                                                     double& tg, double zthr) {
    cd = 1.0; sd = 0.0; tg = 0.0;
    const double g2 = g * g;
    if (g2 > 0.0 && aa > zthr && bb > zthr && g2 > tol2 * aa * bb) {
        const double d = 0.5 * (bb - aa);
        const double ag = fabs(g), ad = fabs(d);
        const double mx = ag > ad ? ag : ad;
        const int ex = (__double2hiint(mx) >> 20) & 0x7ff;
        const double sc = __hiloint2double((2046 - ex) << 20, 0);     // mx * sc in [1, 2)
        const float fd = (float)(ad * sc), fg = (float)(ag * sc);
        const float x = fmaf(fd, fd, fg * fg);
        const float fu = fmaf(x, rsqrtf(x), fd);                      // |d| + sqrt(d^2 + g^2)
        const float rho = rsqrtf(fmaf(fu, fu, fg * fg));
        const float ft = __fdividef(fg, fu);                          // tan of the rotation angle (steering only)
        double c0 = (double)(fu * rho), s0 = (double)(fg * rho);
        if (fg == 0.0f) { c0 = 1.0; s0 = ag / (2.0 * ad); }           // |g| / |d| below fp32 range: tiny angle
        const bool neg = (d >= 0.0) != (g >= 0.0);
        cd = c0;
        sd = neg ? -s0 : s0;
        const double tga = (double)ft * ag;
        tg = d >= 0.0 ? tga : -tga;
    }
}
// (c, s) = (cd, sd) (1 - e/2 + 3 e^2/8): orthogonal to ~e^3
__device__ __forceinline__ void jacobi_rotation_polish(double cd, double sd, double& c, double& s) {
    const double e = fma(cd, cd, fma(sd, sd, -1.0));
    const double corr = fma(e, fma(e, 0.375, -0.5), 1.0);
    c = cd * corr; s = sd * corr;
}

template <int MODE = 0>   // 0: product; 1: rotation warp only, 2: apply warps only, 3: chain skeleton (tools/probe_eig.cu)
__device__ __forceinline__ double* jacobi_eig32_v2_real(double* Gs, double* Gs2, double* Ws, double* scratch, double* rot_c,
                                                        double* rot_s, double tol_rot, int tid, int inner, double zthr) {
    double* Gcur = Gs;
    double* Gnxt = Gs2;
    const int nrr = 16 * inner;
    const double tol2 = tol_rot * tol_rot;
    const int lane = tid & 31, warp = tid >> 5;
    double c = 1.0, sn = 0.0, tg = 0.0, aa = 0.0, bb = 0.0;
    if (warp == 0 && lane < 16) {
        const int pp = lane, qq = 16 + lane;
        aa = Gcur[pp * GP + pp]; bb = Gcur[qq * GP + qq];
        jacobi_rotation_seed(aa, bb, Gcur[pp * GP + qq], tol2, c, sn, tg, zthr);
        rot_c[lane] = c; rot_s[lane] = sn;
    }
    // G-apply threads (warps 1..5): block (ta <= tb) of the upper block triangle, task index u = tid - 32 in 0..135
    int ta = 0, tb = 0;
    const int u = tid - 32;
    if (u >= 0 && u < 136) {
        int t = 0;
#pragma unroll
        for (int k = 1; k < 16; ++k) if (u >= 16 * k - k * (k - 1) / 2) t = k;
        ta = t; tb = ta + (u - (16 * t - t * (t - 1) / 2));
    }
    // W warp: row `lane` of W = I in registers, wp[k] = W[lane][k], wq[k] = W[lane][16 + ((k + step) & 15)]
    double wp[16], wq[16];
    if (warp == 7) {
#pragma unroll
        for (int k = 0; k < 16; ++k) { wp[k] = lane == k ? 1.0 : 0.0; wq[k] = lane == 16 + k ? 1.0 : 0.0; }
    }
    __syncthreads();
    long long busy = 0;    // MODE 4 (probe): cycles between leaving a step barrier and arriving at the next one
    for (int rr = 0; rr < nrr; ++rr) {
        const int cur = (rr & 1) * 16, nxt = 16 - cur;
        long long tb0 = 0;
        if (MODE == 4) tb0 = clock64();
        if (warp == 0) {
            if (MODE != 2 && lane < 16 && rr + 1 < nrr) {
                // pivot of step rr + 1 for pair (p, q'), q' = partner of pair p + 1 at step rr: G'[p][q'] from four
                // entries of G and the two rotations involved (seed accuracy: it only steers)
                const int nb = (lane + 1) & 15;
                const int pp = lane, qq = 16 + ((lane + rr) & 15);
                const int pn = nb, qn = 16 + ((nb + rr) & 15);
                const double gpp = Gcur[pp * GP + pn], gpq = Gcur[pp * GP + qn];
                const double gqp = Gcur[qq * GP + pn], gqq = Gcur[qq * GP + qn];
                const double cn = __shfl_sync(0x0000ffffu, c, nb), snn = __shfl_sync(0x0000ffffu, sn, nb);
                const double bbn = __shfl_sync(0x0000ffffu, bb + tg, nb);
                const double k0 = c * snn, k1 = c * cn, k2 = sn * snn, k3 = sn * cn;
                const double gnew = fma(gpp, k0, gpq * k1) - fma(gqp, k2, gqq * k3);
                aa = aa - tg; bb = bbn;
                if (MODE == 3) { c = 1.0 - gnew * 1e-300; sn = aa * 1e-300; tg = bb * 1e-300; }   // skeleton only (probe)
                else jacobi_rotation_seed(aa, bb, gnew, tol2, c, sn, tg, zthr);
                rot_c[nxt + lane] = c; rot_s[nxt + lane] = sn;
            }
        } else if (MODE != 1 && MODE != 3 && warp < 6) {
            if (u < 136) {
                const int pa = ta, qa = 16 + ((ta + rr) & 15);
                const int pb = tb, qb = 16 + ((tb + rr) & 15);
                double ca, sa, cb, sb;
                jacobi_rotation_polish(rot_c[cur + ta], rot_s[cur + ta], ca, sa);
                jacobi_rotation_polish(rot_c[cur + tb], rot_s[cur + tb], cb, sb);
                const double g00 = Gcur[pa * GP + pb], g01 = Gcur[pa * GP + qb];
                const double g10 = Gcur[qa * GP + pb], g11 = Gcur[qa * GP + qb];
                // rows: r0 = c g0. - s g1., r1 = s g0. + c g1.
                const double r00 = fma(g00, ca, -(g10 * sa)), r01 = fma(g01, ca, -(g11 * sa));
                const double r10 = fma(g00, sa, g10 * ca), r11 = fma(g01, sa, g11 * ca);
                // columns: (.0, .1) -> (c .0 - s .1, s .0 + c .1)
                const double o00 = fma(r00, cb, -(r01 * sb)), o01 = fma(r00, sb, r01 * cb);
                const double o10 = fma(r10, cb, -(r11 * sb)), o11 = fma(r10, sb, r11 * cb);
                Gnxt[pa * GP + pb] = o00; Gnxt[pa * GP + qb] = o01;
                Gnxt[qa * GP + pb] = o10; Gnxt[qa * GP + qb] = o11;
                if (ta != tb) {
                    Gnxt[pb * GP + pa] = o00; Gnxt[qb * GP + pa] = o01;
                    Gnxt[pb * GP + qa] = o10; Gnxt[qb * GP + qa] = o11;
                }
            }
        } else if (MODE != 1 && MODE != 3 && warp == 7) {
            // lanes 0..15 polish the 16 rotations of the step once, every lane reads them back (broadcast loads)
            if (lane < 16) {
                double cc, ss;
                jacobi_rotation_polish(rot_c[cur + lane], rot_s[cur + lane], cc, ss);
                scratch[lane] = cc; scratch[16 + lane] = ss;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const double cc = scratch[k], ss = scratch[16 + k];
                const double a = wp[k], b = wq[k];
                wp[k] = fma(a, cc, -(b * ss));
                wq[k] = fma(a, ss, b * cc);
            }
            if (rr + 1 < nrr) {
                // next step pairs p with the q-column one further on: rotate the q-half by one register
                const double w0 = wq[0];
#pragma unroll
                for (int k = 0; k < 15; ++k) wq[k] = wq[k + 1];
                wq[15] = w0;
            } else {
                // last step (rr = 16 inner - 1): wq[k] holds column 16 + ((k + 15) & 15)
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    Ws[k * WP + lane] = wp[k];
                    Ws[(16 + ((k + 15) & 15)) * WP + lane] = wq[k];
                }
            }
            __syncwarp();
        }
        if (MODE == 4) busy += clock64() - tb0;
        __syncthreads();
        if (MODE != 1 && MODE != 3) { double* tmp = Gcur; Gcur = Gnxt; Gnxt = tmp; }
    }
    if (MODE == 4 && lane == 0) scratch[32 + warp] = (double)busy;
    return Gcur;
}

template <bool CPLX>
__device__ __forceinline__ typename Sc<CPLX>::T* jacobi_eig32_v2(
    typename Sc<CPLX>::T* Gs, typename Sc<CPLX>::T* Gs2, typename Sc<CPLX>::T* Ws, typename Sc<CPLX>::T* rot_ph,
    double* rot_c, double* rot_s, double tol_rot, int tid, int inner, double zthr) {
    if constexpr (CPLX) return jacobi_eig32_pipelined<true>(Gs, Gs2, Ws, rot_ph, rot_c, rot_s, tol_rot, tid, inner, zthr);
    else return jacobi_eig32_v2_real<0>(Gs, Gs2, Ws, rot_ph, rot_c, rot_s, tol_rot, tid, inner, zthr);
}

}  // namespace
}  // namespace dla
}  // namespace t4b
