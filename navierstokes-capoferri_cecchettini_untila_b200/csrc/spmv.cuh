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

template <int L>
__device__ __forceinline__ double row_dot(const CsrView &A, int64_t row, const double *__restrict__ x, int sub) {
  const int64_t b = __ldg(A.rowptr + row), e = __ldg(A.rowptr + row + 1);
  double s = 0;
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
__global__ void __launch_bounds__(256) spmv_kernel(CsrView A, const double *__restrict__ x,
                                                   const double *__restrict__ w, const double *__restrict__ d,
                                                   double *__restrict__ y) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  const int sub = threadIdx.x % L;
  const bool live = g < A.n_rows;
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
__global__ void __launch_bounds__(256) cheb_sweep_kernel(CsrView M, const double *__restrict__ dinv,
                                                         const double *__restrict__ b, const double *__restrict__ z,
                                                         double *__restrict__ d, double *__restrict__ znew, double c1,
                                                         double c2) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  const int sub = threadIdx.x % L;
  const bool live = g < M.n_rows;
  const double s = sub_reduce<L>(live ? row_dot<L>(M, g, z, sub) : 0.0);
  if (sub == 0 && live) {
    const double dn = c1 * d[g] + c2 * dinv[g] * (b[g] - s);
    d[g] = dn;
    znew[g] = z[g] + dn;
  }
}

// ---------------------------------------------------------------------------
// Node-block form of the velocity block: A00 = F_s (x) I_dim, stored as the
// scalar node-level CSR F_s.  Per stored non-zero 12 B are streamed for 2*dim
// flops instead of 12*dim^2 B in the canonical block.
//
// The vector that is GATHERED is kept in a padded node layout [node][PAD],
// PAD = 4 in 3D (2 in 2D), so that the dim values of a neighbour node sit in
// one aligned 32-byte sector and are fetched by one 128-bit + one 64-bit load.
// (ncu on the unpadded version: 1.75 sectors and 3 load instructions per
// non-zero made L1/LSU the busiest unit at 64 %, DRAM at 34 %.)
// ---------------------------------------------------------------------------
template <int DIM>
struct NodePad {
  static constexpr int value = DIM == 3 ? 4 : 2;
};

template <int DIM>
__device__ __forceinline__ void gather_node(const double *__restrict__ xpad, uint32_t col, double (&v)[DIM]) {
  const double2 *p = reinterpret_cast<const double2 *>(xpad + (size_t)NodePad<DIM>::value * col);
  const double2 a = __ldg(p);
  v[0] = a.x;
  v[1] = a.y;
  if constexpr (DIM == 3) v[2] = __ldg(reinterpret_cast<const double *>(p + 1));
}

template <int DIM, int L>
__device__ __forceinline__ void node_row_dot(const CsrView &F, int64_t node, const double *__restrict__ xpad, int sub,
                                             double (&s)[DIM]) {
  const int64_t b = __ldg(F.rowptr + node), e = __ldg(F.rowptr + node + 1);
  double t[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) t[c] = 0.0;
  int64_t k = b + sub;
  // four entries per lane in flight: all (value, column) loads first, then the four gathers
  for (; k + 3 * L < e; k += 4 * L) {
    const double v0 = ld_stream(F.val + k), v1 = ld_stream(F.val + k + L);
    const double v2 = ld_stream(F.val + k + 2 * L), v3 = ld_stream(F.val + k + 3 * L);
    const uint32_t c0 = ld_stream(F.colind + k), c1 = ld_stream(F.colind + k + L);
    const uint32_t c2 = ld_stream(F.colind + k + 2 * L), c3 = ld_stream(F.colind + k + 3 * L);
    double x0[DIM], x1[DIM], x2[DIM], x3[DIM];
    gather_node<DIM>(xpad, c0, x0);
    gather_node<DIM>(xpad, c1, x1);
    gather_node<DIM>(xpad, c2, x2);
    gather_node<DIM>(xpad, c3, x3);
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      s[c] += v0 * x0[c];
      t[c] += v1 * x1[c];
      s[c] += v2 * x2[c];
      t[c] += v3 * x3[c];
    }
  }
  for (; k + L < e; k += 2 * L) {
    const double v0 = ld_stream(F.val + k), v1 = ld_stream(F.val + k + L);
    const uint32_t c0 = ld_stream(F.colind + k), c1 = ld_stream(F.colind + k + L);
    double x0[DIM], x1[DIM];
    gather_node<DIM>(xpad, c0, x0);
    gather_node<DIM>(xpad, c1, x1);
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      s[c] += v0 * x0[c];
      t[c] += v1 * x1[c];
    }
  }
  if (k < e) {
    const double v0 = ld_stream(F.val + k);
    double x0[DIM];
    gather_node<DIM>(xpad, ld_stream(F.colind + k), x0);
#pragma unroll
    for (int c = 0; c < DIM; ++c) s[c] += v0 * x0[c];
  }
#pragma unroll
  for (int c = 0; c < DIM; ++c) s[c] += t[c];
}

// standard [node][dim] -> padded [node][PAD] (all local nodes incl. ghosts)
template <int DIM>
__global__ void pad_nodes_kernel(int64_t n_nodes, const double *__restrict__ x, double *__restrict__ xpad) {
  constexpr int PAD = NodePad<DIM>::value;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes * PAD) return;
  const int64_t node = t / PAD;
  const int c = (int)(t % PAD);
  xpad[t] = c < DIM ? x[node * DIM + c] : 0.0;
}

// velocity rows of the block product:  y_u = F x_u (+ A01 x_p when a01.rowptr)
//   MODE 0: y = ..., MODE 3: y = d .* (F x)   (power iteration on D^-1 F)
template <int DIM, int L, int MODE>
__global__ void __launch_bounds__(256) fs_apply_kernel(CsrView F, CsrView a01, const double *__restrict__ xpad,
                                                       const double *__restrict__ xp, const double *__restrict__ d,
                                                       double *__restrict__ y) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  const int sub = threadIdx.x % L;
  const bool live = g < F.n_rows;
  double s[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) s[c] = 0.0;
  if (live) {
    node_row_dot<DIM, L>(F, g, xpad, sub, s);
    if (a01.rowptr != nullptr) {
#pragma unroll
      for (int c = 0; c < DIM; ++c) s[c] += row_dot<L>(a01, (int64_t)DIM * g + c, xp, sub);
    }
  }
#pragma unroll
  for (int c = 0; c < DIM; ++c) s[c] = sub_reduce<L>(s[c]);
  if (live && sub < DIM) {
    double sc = s[0];
#pragma unroll
    for (int c = 1; c < DIM; ++c)
      if (sub == c) sc = s[c];
    const int64_t i = (int64_t)DIM * g + sub;
    y[i] = MODE == 3 ? d[i] * sc : sc;
  }
}

// Chebyshev-Jacobi sweep on F (see cheb_sweep_kernel), node-block form.  z is
// padded; the result goes to a padded buffer (next sweep) or, for the last
// sweep, to a standard-layout vector.
template <int DIM, int L, bool OUT_STD>
__global__ void __launch_bounds__(256) fs_cheb_sweep_kernel(CsrView F, const double *__restrict__ dinv,
                                                            const double *__restrict__ b,
                                                            const double *__restrict__ zpad, double *__restrict__ d,
                                                            double *__restrict__ znew, double c1, double c2) {
  constexpr int PAD = NodePad<DIM>::value;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  const int sub = threadIdx.x % L;
  const bool live = g < F.n_rows;
  double s[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) s[c] = 0.0;
  if (live) node_row_dot<DIM, L>(F, g, zpad, sub, s);
#pragma unroll
  for (int c = 0; c < DIM; ++c) s[c] = sub_reduce<L>(s[c]);
  if (live && sub < DIM) {
    // lane c of the sub-warp finishes component c
    double sc = s[0];
#pragma unroll
    for (int c = 1; c < DIM; ++c)
      if (sub == c) sc = s[c];
    const int64_t i = (int64_t)DIM * g + sub;
    const double dn = c1 * d[i] + c2 * dinv[i] * (b[i] - sc);
    d[i] = dn;
    const double zn = zpad[(int64_t)PAD * g + sub] + dn;
    if (OUT_STD)
      znew[i] = zn;
    else
      znew[(int64_t)PAD * g + sub] = zn;
  }
}

// first sweep on F with zero initial guess: d = Dinv .* b / theta (standard), z = d (padded)
template <int DIM>
__global__ void fs_cheb_first_kernel(int64_t n_u, const double *__restrict__ dinv, const double *__restrict__ b,
                                     double inv_theta, double *__restrict__ d, double *__restrict__ zpad) {
  constexpr int PAD = NodePad<DIM>::value;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_u) {
    const double v = dinv[i] * b[i] * inv_theta;
    d[i] = v;
    zpad[(i / DIM) * PAD + i % DIM] = v;
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
// of outer products over the velocity rows this rank owns:
//   S[V,W] += B[V,u] Di[u] Bt[u,W],  B[V,u] = a10t[u,V] (pattern of row u of A01).
// One warp per velocity dof; lanes stride over the (V,W) pairs of the row.  On
// several GPUs every rank adds its owned rows and the values are all-reduced.
__global__ void __launch_bounds__(256) schur_outer_kernel(CsrView Bt, const double *__restrict__ a10t,
                                                          const double *__restrict__ di, CsrView S) {
  const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (u >= Bt.n_rows) return;
  const int64_t b = Bt.rowptr[u];
  const int len = (int)(Bt.rowptr[u + 1] - b);
  const double d = di[u];
  for (int p = lane; p < len * len; p += 32) {
    const int i = p / len, j = p % len;
    const double t = Bt.val[b + j];
    if (t == 0.0) continue;  // constrained rows of Bt are zero
    const uint32_t V = Bt.colind[b + i], W = Bt.colind[b + j];
    int64_t lo = S.rowptr[V], hi = S.rowptr[V + 1];
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (S.colind[mid] < W)
        lo = mid + 1;
      else
        hi = mid;
    }
    atomicAdd(S.val + lo, a10t[b + i] * d * t);
  }
}

// halo pack: buf[i][c] = x[width*idx[i] + c], width = dim (standard) or PAD (padded vectors)
__global__ void halo_pack_kernel(int64_t n, int width, const uint32_t *__restrict__ idx, const double *__restrict__ x,
                                 double *__restrict__ buf) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n * width) buf[t] = x[(int64_t)width * idx[t / width] + t % width];
}

}  // namespace nsb
