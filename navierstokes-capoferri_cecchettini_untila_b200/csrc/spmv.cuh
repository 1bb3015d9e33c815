// CSR sparse matrix-vector kernels on the canonical (reference) block pattern.
//
// Replaces TrilinosWrappers::BlockSparseMatrix::vmult / SparseMatrix::vmult as
// used inside SolverGMRES and PreconditionASIMPLE::vmult (reference
// src/NavierStokes.cpp:377, 982, 992).  Everything here is HBM-bound: per
// non-zero 8 B value + 4 B column index are streamed once (ld.global.nc,
// evict-first), x is gathered through L2.  Rows are long (A00 ~81, A10 ~169,
// S ~53 non-zeros in 3D), so a sub-warp of L lanes walks one row with
// coalesced 8*L-byte segments and finishes with a shuffle reduction.
#pragma once
#include "common.cuh"

namespace nsb {

__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p) { return __ldcs(p); }

// partial dot product of one CSR row with x over the lanes of a sub-warp.
// Entries are taken in batches of kBatch per lane: all column/value loads of a
// batch are issued first, then all gathers, so that every lane keeps kBatch
// independent dependent-load chains in flight (ncu: 91 % of the stall cycles of
// the two-deep version were long-scoreboard waits, 0.28 eligible warps/scheduler).
constexpr int kBatch = 4;
// Requesting the batches after the first into L2 (prefetch.global.L2) as soon as the row bounds are known -- the
// idea that pays in the slab kernels (slab.cuh) -- does NOT pay here: measured on B200 at 9.7 M DoFs, A10 product
// 0.181 -> 0.202 ms, sweep on S 0.0785 -> 0.0834 ms (rows of 169 / 53 entries are two batches long; the extra
// requests cost more than the one L2-latency they save).  Off.
#ifndef NSB_CSR_PF
#define NSB_CSR_PF 0
#endif
__device__ __forceinline__ void csr_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int L>
__device__ __forceinline__ double row_dot(const CsrView &A, int64_t row, const double *__restrict__ x, int sub) {
  const int64_t b = __ldg(A.rowptr + row), e = __ldg(A.rowptr + row + 1);
  double s = 0;
  if (NSB_CSR_PF) {
    for (int64_t kk = b + kBatch * L + 16 * sub; kk < e; kk += 16 * L) csr_prefetch_l2(A.val + kk);
    for (int64_t kk = b + kBatch * L + 32 * sub; kk < e; kk += 32 * L) csr_prefetch_l2(A.colind + kk);
  }
  for (int64_t k = b + sub; k < e; k += kBatch * L) {
    double v[kBatch];
    uint32_t c[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int64_t kk = k + j * L;
      const bool in = kk < e;
      v[j] = in ? ld_stream(A.val + kk) : 0.0;
      c[j] = in ? ld_stream(A.colind + kk) : 0u;
    }
    double xv[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) xv[j] = __ldg(x + c[j]);
#pragma unroll
    for (int j = 0; j < kBatch; ++j) s += v[j] * xv[j];
  }
  return s;
}

template <int L>
__device__ __forceinline__ double sub_reduce(double s) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// y = [A00 x_u + A01 x_p ; A10 x_u]   (A11 is empty)
template <int L>
__global__ void __launch_bounds__(256) block_spmv_kernel(CsrView a00, CsrView a01, CsrView a10,
                                                         const double *__restrict__ x, double *__restrict__ y) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  const int sub = threadIdx.x % L;
  const int64_t n_u = a00.n_rows, N = n_u + a10.n_rows;
  double s = 0;  // lanes past the end stay for the full-mask shuffles
  if (g < n_u)
    s = row_dot<L>(a00, g, x, sub) + row_dot<L>(a01, g, x + n_u, sub);
  else if (g < N)
    s = row_dot<L>(a10, g - n_u, x, sub);
  s = sub_reduce<L>(s);
  if (sub == 0 && g < N) y[g] = s;
}

// Generic single-block product with a fused epilogue:
//   mode 0: y = A x
//   mode 1: y = w - A x                     (vec1 = src1 - B vec0, reference :982-983)
//   mode 2: y = w - d .* (A x)              (dst0 = vec0 - Di .* (Bt dst1), reference :992-994)
//   mode 3: y = d .* (A x)                  (power iteration on D^-1 A)
template <int L, int MODE>
__global__ void __launch_bounds__(256) spmv_kernel(CsrView A, int64_t row0, int64_t n_loc, const double *__restrict__ x,
                                                   const double *__restrict__ w, const double *__restrict__ d,
                                                   double *__restrict__ y) {
  // rows [row0, row0 + n_loc) of A (a rank's owned rows of a replicated matrix; the whole matrix otherwise)
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L + row0;
  const int sub = threadIdx.x % L;
  const bool live = g < row0 + n_loc;
  const double s = sub_reduce<L>(live ? row_dot<L>(A, g, x, sub) : 0.0);
  if (sub == 0 && live) {
    if (MODE == 0) y[g] = s;
    if (MODE == 1) y[g] = w[g] - s;
    if (MODE == 2) y[g] = w[g] - d[g] * s;
    if (MODE == 3) y[g] = d[g] * s;
  }
}

// One Chebyshev-Jacobi sweep for M z = b (the Jacobi-type inner sweep that
// replaces ILU + inner GMRES, reference :978-981, 986-989):
//   dnew = c1 * d + c2 * Dinv .* (b - M z);  znew = z + dnew
// z and znew are distinct buffers (the product gathers z).
template <int L>
__global__ void __launch_bounds__(256) cheb_sweep_kernel(CsrView M, int64_t row0, int64_t n_loc,
                                                         const double *__restrict__ dinv, const double *__restrict__ b,
                                                         const double *__restrict__ z, double *__restrict__ d,
                                                         double *__restrict__ znew, double c1, double c2) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L + row0;
  const int sub = threadIdx.x % L;
  const bool live = g < row0 + n_loc;
  const double s = sub_reduce<L>(live ? row_dot<L>(M, g, z, sub) : 0.0);
  if (sub == 0 && live) {
    const double dn = c1 * d[g] + c2 * dinv[g] * (b[g] - s);
    d[g] = dn;
    znew[g] = z[g] + dn;
  }
}

// first sweep with zero initial guess: d = z = Dinv .* b / theta
__global__ void cheb_first_kernel(int64_t n, const double *__restrict__ dinv, const double *__restrict__ b,
                                  double inv_theta, double *__restrict__ d, double *__restrict__ z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double v = dinv[i] * b[i] * inv_theta;
    d[i] = v;
    z[i] = v;
  }
}

// dinv[i] = 1 / M[i / rep, i / rep]  (rep = dim for the node-block F, 1 for S)
__global__ void diag_inverse_kernel(int64_t n, int rep, const double *__restrict__ val,
                                    const int64_t *__restrict__ diagpos, double *__restrict__ dinv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dinv[i] = 1.0 / val[diagpos[i / rep]];
}

// S = B diag(Di) Bt on the precomputed pattern of S (reference :956), as a sum
// of outer products over the velocity NODES this rank owns: the dim rows of a node
// share their pattern in A01 and their Di, so
//   S[V,W] += Di[a] sum_c B[V,(a,c)] Bt[(a,c),W],  B[V,u] = a10t[u,V] (pattern of row u of A01),
// one atomic per (V,W) pair and node instead of per dof.  One warp per node; lanes
// stride over the (V,W) pairs of the row.  On several GPUs every rank adds its
// owned rows and the values are all-reduced.
template <int DIM>
__global__ void __launch_bounds__(256) schur_outer_kernel(CsrView Bt, const double *__restrict__ a10t,
                                                          const double *__restrict__ di, CsrView S) {
  const int64_t a = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (DIM * a >= Bt.n_rows) return;
  int64_t b[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) b[c] = Bt.rowptr[DIM * a + c];
  const int len = (int)(Bt.rowptr[DIM * a + 1] - b[0]);
  const double d = di[DIM * a];
  for (int p = lane; p < len * len; p += 32) {
    const int i = p / len, j = p % len;
    double t = 0.0;
#pragma unroll
    for (int c = 0; c < DIM; ++c) t += a10t[b[c] + i] * Bt.val[b[c] + j];
    if (t == 0.0) continue;  // constrained rows of Bt are zero
    const uint32_t V = Bt.colind[b[0] + i], W = Bt.colind[b[0] + j];
    int64_t lo = S.rowptr[V], hi = S.rowptr[V + 1];
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (S.colind[mid] < W)
        lo = mid + 1;
      else
        hi = mid;
    }
    atomicAdd(S.val + lo, d * t);
  }
}

// halo pack: buf[i][c] = x[width*idx[i] + c], width = dim
__global__ void halo_pack_kernel(int64_t n, int width, const uint32_t *__restrict__ idx, const double *__restrict__ x,
                                 double *__restrict__ buf) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n * width) buf[t] = x[(int64_t)width * idx[t / width] + t % width];
}

}  // namespace nsb
