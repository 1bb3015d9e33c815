// Multilevel Jacobi-type solve for the Schur approximation S = B diag(Di) Bt
// (the "inner loop B" of PreconditionASIMPLE::vmult, reference
// src/NavierStokes.cpp:986-989, where the reference runs ILU(0)-preconditioned
// GMRES to 1e-2).
//
// S is symmetric positive definite and Laplacian-like on the pressure mesh, so
// a single-level Chebyshev-Jacobi polynomial needs ~sqrt(cond) ~ 1/h sweeps
// (SURVEY.md H4).  Here the same Chebyshev-Jacobi sweeps are used as the
// smoother of a V-cycle over an aggregation hierarchy:
//   * aggregates: greedy on a strength-of-connection graph (see coarsen for the measures),
//     at most `max_agg` members, built ONCE on the host from the first S;
//   * prolongation: piecewise constant, coarse correction scaled by omega;
//   * coarse operators: Galerkin sums S_c[I,J] = sum_{i in I, j in J} S[i,j],
//     recomputed on the device every step through a fine-nnz -> coarse-nnz map.
// The cycle is a fixed linear operator (no inner reductions, no stale guesses).
// Measured and dropped (B200, 9.7 M DoFs, round 2): the five launches of a coarse level (first sweep, residual,
// restriction, prolongation, post-sweep) as two fused kernels that recompute the pre-smoothed iterate where it is
// gathered -- same iteration counts, same time (406 vs 403 ms per solve): the coarse chain is bound by the
// dependent latencies of its small kernels, not by their number.
#pragma once
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace nsb {

// coarse.val[pos[k]] += fine.val[k]
__global__ void galerkin_sum_kernel(int64_t nnz, const double *__restrict__ fine, const int64_t *__restrict__ pos,
                                    double *coarse) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nnz) atomicAdd(coarse + pos[k], fine[k]);
}

// rc[I] = sum of r over the members of aggregate I (fixed order: deterministic)
__global__ void restrict_kernel(int64_t n_coarse, const int64_t *__restrict__ agg_ptr,
                                const uint32_t *__restrict__ agg_idx, const double *__restrict__ r,
                                double *__restrict__ rc) {
  const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= n_coarse) return;
  double s = 0;
  for (int64_t k = agg_ptr[I]; k < agg_ptr[I + 1]; ++k) s += r[agg_idx[k]];
  rc[I] = s;
}

// z[i] += omega * ec[agg[i]]
__global__ void prolong_add_kernel(int64_t n, const uint32_t *__restrict__ agg, const double *__restrict__ ec,
                                   double omega, double *__restrict__ z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[i] += omega * ec[agg[i]];
}

// The coarsest level (<= kCoarseFusedMax rows): all k Chebyshev-Jacobi sweeps of M z = b (zero initial guess,
// same recurrence as cheb_smooth) in ONE single-CTA launch, the iterate in shared memory -- the 16
// separate sweep launches of 3-4 us each were a quarter of the V-cycle's time.  One thread per row.
constexpr int kCoarseFusedMax = 1024;
__global__ void __launch_bounds__(kCoarseFusedMax) coarse_cheb_kernel(CsrView M, const double *__restrict__ dinv,
                                                                      const double *__restrict__ b, int k, double lmax,
                                                                      double ratio, double *__restrict__ z_out) {
  __shared__ double z[kCoarseFusedMax];
  const int i = threadIdx.x;
  const int n = (int)M.n_rows;
  const double lmin = lmax / ratio, theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double di = 0, bi = 0, dd = 0, zz = 0;
  int64_t kb = 0, ke = 0;
  if (i < n) {
    di = dinv[i];
    bi = b[i];
    kb = M.rowptr[i];
    ke = M.rowptr[i + 1];
    dd = di * bi / theta;
    zz = dd;
    z[i] = zz;
  }
  __syncthreads();
  double rho = 1.0 / sigma;
  for (int s = 1; s < k; ++s) {
    const double rho_new = 1.0 / (2.0 * sigma - rho);
    double r = bi;
    if (i < n)
      for (int64_t q = kb; q < ke; ++q) r -= M.val[q] * z[M.colind[q]];
    __syncthreads();
    if (i < n) {
      dd = rho_new * rho * dd + 2.0 * rho_new / delta * di * r;
      zz += dd;
      z[i] = zz;
    }
    __syncthreads();
    rho = rho_new;
  }
  if (i < n) z_out[i] = zz;
}

// ---- host-side setup (runs once) -------------------------------------------
struct HostCsr {
  int64_t n = 0;
  std::vector<int64_t> rowptr;
  std::vector<uint32_t> colind;
  std::vector<double> val;
};

struct HostCoarsening {
  std::vector<uint32_t> agg;       // fine -> coarse
  std::vector<int64_t> agg_ptr;    // coarse -> members
  std::vector<uint32_t> agg_idx;
  std::vector<int64_t> pos;        // fine nnz -> coarse nnz
  HostCsr coarse;
};

// Greedy aggregation on the strength graph, visiting rows in index order; a
// root takes its (up to max_agg-1) strongest still-free strong neighbours; a
// row whose strong neighbours are all taken joins its strongest neighbour.
// Strength of connection, three measures (nsb_set_schur_strength; defaults per dimension in amg_build):
//   0 (Ruge-Stueben): j is strong for i when -s_ij >= theta * max_k(-s_ik) -- negative couplings only, relative to
//     the row.  S = B D^-1 Bt of Taylor-Hood has a fifth of its off-diagonal entries POSITIVE (aggregating along them
//     puts nodes together whose smooth-error values differ), and an absolute threshold does not transfer across a
//     graded mesh: on the NACA 2408 / 10 degrees mesh of run_test.sh (cells from 0.002 to 0.03, cond(D^-1 S) = 8000 at
//     720 rows) measure 1 at 0.08 leaves V S with eigenvalues down to 0.005 and the outer GMRES needs 750 iterations
//     where an exact Schur solve needs 129; measure 0 at 0.35: 122 (tests/amg_emul.py; B200: 121).
//   1: |s_ij| >= theta sqrt(s_ii s_jj), theta halved per level.  On the uniform 3D cylinder mesh at 9.7 M DoFs (B200)
//     it needs 99 outer iterations where measure 0 needs 141 (theta 0.35) / 107 (0.2, halved per level).
// measure 0: relative, negative couplings (above);  1: |s_ij| >= theta sqrt(s_ii s_jj) (absolute, symmetric);
//         2: |s_ij| >= theta max_k |s_ik| (relative, either sign)
inline HostCoarsening coarsen(const HostCsr &M, double theta, int max_agg, const int32_t *owner = nullptr,
                              int measure = 0) {
  const int64_t n = M.n;
  HostCoarsening C;
  std::vector<double> rowmax(n, 0.0), diag(n, 0.0);  // strongest coupling of the row (all neighbours, any owner)
  for (int64_t i = 0; i < n; ++i)
    for (int64_t k = M.rowptr[i]; k < M.rowptr[i + 1]; ++k) {
      if (M.colind[k] != (uint32_t)i)
        rowmax[i] = std::max(rowmax[i], measure == 0 ? -M.val[k] : std::fabs(M.val[k]));
      else
        diag[i] = std::fabs(M.val[k]);
    }
  C.agg.assign(n, UINT32_MAX);
  uint32_t na = 0;
  std::vector<std::pair<double, uint32_t>> nb;
  for (int64_t i = 0; i < n; ++i) {
    if (C.agg[i] != UINT32_MAX) continue;
    nb.clear();
    double best_w = -1;
    uint32_t best = UINT32_MAX;
    for (int64_t k = M.rowptr[i]; k < M.rowptr[i + 1]; ++k) {
      const uint32_t j = M.colind[k];
      if (j == (uint32_t)i) continue;
      if (owner && owner[j] != owner[i]) continue;
      const double w = measure == 0 ? -M.val[k] : std::fabs(M.val[k]);
      if (!(w > 0) || w < theta * (measure == 1 ? std::sqrt(diag[i] * diag[j]) : rowmax[i])) continue;
      if (C.agg[j] == UINT32_MAX)
        nb.push_back({w, j});
      else if (w > best_w) {
        best_w = w;
        best = j;
      }
    }
    if (nb.empty() && best != UINT32_MAX) {
      C.agg[i] = C.agg[best];
      continue;
    }
    std::stable_sort(nb.begin(), nb.end(), [](const auto &a, const auto &b) { return a.first > b.first; });
    C.agg[i] = na;
    for (size_t t = 0; t < nb.size() && (int)t < max_agg - 1; ++t) C.agg[nb[t].second] = na;
    ++na;
  }
  // members
  C.agg_ptr.assign((size_t)na + 1, 0);
  for (int64_t i = 0; i < n; ++i) ++C.agg_ptr[C.agg[i] + 1];
  for (uint32_t I = 0; I < na; ++I) C.agg_ptr[I + 1] += C.agg_ptr[I];
  C.agg_idx.resize(n);
  {
    std::vector<int64_t> fill(C.agg_ptr.begin(), C.agg_ptr.end() - 1);
    for (int64_t i = 0; i < n; ++i) C.agg_idx[fill[C.agg[i]]++] = (uint32_t)i;
  }
  // coarse pattern and the fine-nnz -> coarse-nnz map
  HostCsr &A = C.coarse;
  A.n = na;
  A.rowptr.assign((size_t)na + 1, 0);
  std::vector<std::vector<uint32_t>> rows(na);
  for (uint32_t I = 0; I < na; ++I) {
    auto &row = rows[I];
    for (int64_t m = C.agg_ptr[I]; m < C.agg_ptr[I + 1]; ++m) {
      const int64_t i = C.agg_idx[m];
      for (int64_t k = M.rowptr[i]; k < M.rowptr[i + 1]; ++k) row.push_back(C.agg[M.colind[k]]);
    }
    std::sort(row.begin(), row.end());
    row.erase(std::unique(row.begin(), row.end()), row.end());
    A.rowptr[I + 1] = A.rowptr[I] + (int64_t)row.size();
  }
  A.colind.resize((size_t)A.rowptr[na]);
  for (uint32_t I = 0; I < na; ++I) std::copy(rows[I].begin(), rows[I].end(), A.colind.begin() + A.rowptr[I]);
  A.val.assign(A.colind.size(), 0.0);
  C.pos.resize(M.colind.size());
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t I = C.agg[i];
    const auto b = A.colind.begin() + A.rowptr[I], e = A.colind.begin() + A.rowptr[I + 1];
    for (int64_t k = M.rowptr[i]; k < M.rowptr[i + 1]; ++k) {
      const int64_t p = std::lower_bound(b, e, C.agg[M.colind[k]]) - A.colind.begin();
      C.pos[k] = p;
      A.val[p] += M.val[k];
    }
  }
  return C;
}

// one level on the device (level 0 aliases the context's S)
struct AmgLevel {
  int64_t n = 0;
  CsrDev M;                    // unused on level 0
  DevBuf<double> dinv;         // unused on level 0
  DevBuf<int64_t> diag;        // unused on level 0
  DevBuf<double> b, z0, z1, d, r, eig;
  DevBuf<uint32_t> agg, agg_idx;   // to the next level
  DevBuf<int64_t> agg_ptr, pos;
  double lmax = 0;
  bool eig_warm = false;
};

}  // namespace nsb
