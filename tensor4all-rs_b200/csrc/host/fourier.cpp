// Quantics Fourier MPO built on the device side of the seam: the (K+1) x 2 x 2 x (K+1) Chen & Lindsey
// core is closed-form host arithmetic (a few thousand values), the R site tensors are uploaded once and the LU
// compression - the only part with real work - runs through the device simplett path.  Mirrors
// quantics_fourier_mpo (reference crates/tensor4all-quanticstransform/src/fourier.rs:291-404; chebyshev_grid
// :405-431, lagrange_polynomial :432-444, build_dft_core_tensor :445-481).
#include <cmath>
#include <vector>

#include "simplett.h"

namespace t4b {
namespace stt {

Train fourier_mpo(dla::Ctx* c, int r, int k, double sign, double tolerance, std::optional<int64_t> max_bond_dim,
                  bool normalize) {
    T4B_REQUIRE(r >= 2, "Number of sites must be at least 2");
    T4B_REQUIRE(k >= 1 && k <= 1024, "Fourier interpolation order out of range");
    const int n = k + 1;
    const double pi = 3.14159265358979323846;
    std::vector<double> grid(n), w(n);
    for (int j = 0; j < n; ++j) grid[j] = 0.5 * (1.0 - std::cos(pi * (double)j / (double)k));
    for (int j = 0; j < n; ++j) {
        double weight = 1.0;
        for (int m = 0; m < n; ++m)
            if (j != m) weight /= grid[j] - grid[m];
        w[j] = weight;
    }
    auto lagrange = [&](int alpha, double x) {
        if (std::fabs(x - grid[alpha]) < 1e-14) return 1.0;
        double prod = 1.0;
        for (double g : grid) prod *= x - g;
        return prod * w[alpha] / (x - grid[alpha]);
    };
    // core[alpha, tau, sigma, beta] = P_alpha(x) exp(2 pi i sign x tau), x = (sigma + grid[beta]) / 2
    auto core = [&](int alpha, int tau, int sigma, int beta, double* re, double* im) {
        const double x = ((double)sigma + grid[beta]) / 2.0;
        const double p = lagrange(alpha, x);
        const double ph = 2.0 * pi * sign * x * (double)tau;
        *re = p * std::cos(ph);
        *im = p * std::sin(ph);
    };
    Train tt;
    tt.dt = C64;
    tt.rank = 3;
    auto make_site = [&](int64_t l, int64_t rr, const std::vector<double>& host) {
        Site s;
        s.d[0] = l; s.d[1] = 4; s.d[2] = rr;
        s.buf = std::make_shared<Buffer>(c, host.size() * sizeof(double));
        dla::h2d(c, s.buf->p, host.data(), host.size() * sizeof(double));
        dla::sync(c);   // host is a temporary
        return s;
    };
    // Tensor3 [left, s, right], s = tau * 2 + sigma, column-major, interleaved complex
    {   // first: sum over alpha
        std::vector<double> h((size_t)1 * 4 * n * 2, 0.0);
        for (int tau = 0; tau < 2; ++tau)
            for (int sigma = 0; sigma < 2; ++sigma)
                for (int beta = 0; beta < n; ++beta) {
                    double sr = 0.0, si = 0.0;
                    for (int alpha = 0; alpha < n; ++alpha) { double a, b; core(alpha, tau, sigma, beta, &a, &b); sr += a; si += b; }
                    const size_t e = (size_t)(tau * 2 + sigma) + 4 * (size_t)beta;
                    h[2 * e] = sr; h[2 * e + 1] = si;
                }
        tt.sites.push_back(make_site(1, n, h));
    }
    {
        std::vector<double> h((size_t)n * 4 * n * 2, 0.0);
        for (int alpha = 0; alpha < n; ++alpha)
            for (int tau = 0; tau < 2; ++tau)
                for (int sigma = 0; sigma < 2; ++sigma)
                    for (int beta = 0; beta < n; ++beta) {
                        const size_t e = (size_t)alpha + (size_t)n * ((size_t)(tau * 2 + sigma) + 4 * (size_t)beta);
                        core(alpha, tau, sigma, beta, &h[2 * e], &h[2 * e + 1]);
                    }
        for (int i = 1; i + 1 < r; ++i) tt.sites.push_back(make_site(n, n, h));
    }
    {   // last: beta = 0
        std::vector<double> h((size_t)n * 4 * 2, 0.0);
        for (int alpha = 0; alpha < n; ++alpha)
            for (int tau = 0; tau < 2; ++tau)
                for (int sigma = 0; sigma < 2; ++sigma) {
                    const size_t e = (size_t)alpha + (size_t)n * (size_t)(tau * 2 + sigma);
                    core(alpha, tau, sigma, 0, &h[2 * e], &h[2 * e + 1]);
                }
        tt.sites.push_back(make_site(n, 1, h));
    }
    CompressionOptions o;
    o.method = CompressionMethod::LU;
    o.tolerance = tolerance;
    o.max_bond_dim = max_bond_dim;
    o.normalize_error = true;
    compress(c, tt, o);
    if (normalize)
        for (auto& s : tt.sites) dla::scal(c, C64, s.d[0] * s.d[1] * s.d[2], s.buf->p, 1.0 / std::sqrt(2.0));
    return tt;
}

}  // namespace stt
}  // namespace t4b
