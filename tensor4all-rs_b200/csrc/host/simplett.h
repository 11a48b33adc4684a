// Positional tensor-train stack: mirror of tensor4all-simplett's hot functions
//   compress          reference crates/tensor4all-simplett/src/compression.rs:165-501
//   mpo::contract_zipup   reference src/mpo/contract_zipup.rs:45-164
//   mpo::contract_naive / compress_mpo / right_canonicalize
//                     reference src/mpo/contract_naive.rs:41-169, src/mpo/canonical.rs:35-86
//   mpo::factorize_svd    reference src/mpo/factorize.rs:182-302
//   inner_product     reference src/contraction.rs:82-167
// Site tensors are dense column-major device buffers: Tensor3 [left, site, right],
// Tensor4 [left, s1, s2, right] (reference src/types.rs:104-127, src/mpo/types.rs:59-68).
#pragma once
#include <optional>
#include <vector>

#include "factorize.h"
#include "luci.h"

namespace t4b {
namespace stt {

struct Site {
    std::shared_ptr<Buffer> buf;
    int64_t d[4] = {1, 1, 1, 1};   // Tensor3: d[0..2]; Tensor4: d[0..3]
};
struct Train {
    DType dt = F64;
    int rank = 3;   // 3: tensor train, 4: MPO
    std::vector<Site> sites;
};

enum class CompressionMethod { LU, CI, SVD };
struct CompressionOptions {   // reference compression.rs:88-124
    CompressionMethod method = CompressionMethod::LU;
    double tolerance = 1e-12;
    std::optional<int64_t> max_bond_dim;
    bool normalize_error = true;
};
void compress(dla::Ctx*, Train& tt, const CompressionOptions& opts);
// The same two-pass compress for a BATCH of independent tensor trains of equal length (the small-chi regime of the
// north star: C1 "batch 1 and 1024"): at every sweep position the factorisations of all trains are ONE launch of the
// single-CTA SVD kernel and the absorb products ONE ragged batched GEMM launch; the only host synchronisation is one
// read of all spectra per truncating step (the rank rule stays on the host, compression.rs:286-306).  Per train the
// result equals compress() up to rounding.  Falls back to the per-train loop for pivoted methods or matrices that do
// not fit one CTA.
void compress_batched(dla::Ctx*, const std::vector<Train*>& tts, const CompressionOptions& opts);

struct MpoContractionOptions {   // reference mpo/types.rs ContractionOptions
    double tolerance = 1e-12;
    std::optional<int64_t> max_bond_dim;
};
Train contract_zipup(dla::Ctx*, const Train& a, const Train& b, const MpoContractionOptions& opts);
Train contract_naive(dla::Ctx*, const Train& a, const Train& b,
                     const std::optional<MpoContractionOptions>& opts);
void right_canonicalize(dla::Ctx*, Train& mpo);
// bilinear <a,b> (no conjugation, like the reference); returns (re, im)
void inner_product(dla::Ctx*, const Train& a, const Train& b, double* re, double* im);

// quantics_fourier_mpo (reference crates/tensor4all-quanticstransform/src/fourier.rs:291-404): Tensor3 sites
// [left, 4, right] with s = tau*2 + sigma, LU-compressed (FourierOptions: k, sign, tolerance, max_bond_dim, normalize)
Train fourier_mpo(dla::Ctx*, int r, int k, double sign, double tolerance, std::optional<int64_t> max_bond_dim,
                  bool normalize);

// Batched evaluation of a tensor train on the device (TTCache::evaluate_many, reference simplett/src/cache.rs:594-690:
// values = <left environment | right environment>): indices is npts x L point-major on the HOST; out: npts values on
// the device.
void evaluate_many(dla::Ctx*, const Train& tt, int64_t npts, const int64_t* indices, void* out_dev);
// Candidate matrix Pi of the TCI2 two-site update for a TT-valued integrand, built entirely on the device (the batch
// callback of reference tensorci/src/tensorci2.rs:1862-1893 with f = the tensor train): rows i*d_b + s over the ni
// left multi-indices (ni x b, host, point-major), columns s'*nj + j over the nj right multi-indices (nj x (L-b-2));
// Pi[(i,s),(s',j)] = Lenv_i T_b[:, s, :] T_{b+1}[:, s', :] Renv_j as two environment launches and three GEMMs.
// out: (ni*d_b) x (d_{b+1}*nj) device matrix.
void tci2_pi_from_train(dla::Ctx*, const Train& tt, int b, int64_t ni, const int64_t* i_multi, int64_t nj,
                        const int64_t* j_multi, void* pi_dev);

// rank rule shared by compression.rs:286-306 and mpo/factorize.rs:206-250
int64_t simplett_rank(const std::vector<double>& s, double tolerance, bool normalize_error,
                      std::optional<int64_t> max_bond_dim);

}  // namespace stt
}  // namespace t4b
