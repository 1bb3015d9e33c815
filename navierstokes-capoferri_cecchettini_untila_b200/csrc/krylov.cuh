// Fused vector kernels of the GMRES iteration (replaces the l2_norm / dot /
// add / sadd / scale calls deal.II's SolverGMRES and
// PreconditionASIMPLE::vmult issue one at a time, reference
// src/NavierStokes.cpp:348, 377, 978-994).
//
// Orthogonalisation is classical Gram-Schmidt applied twice (CGS2) in three
// passes over the basis instead of four: inner products, then projection fused
// with the second set of inner products, then the second projection fused with
// the norm (ortho_kernel).  All reductions are deterministic: block partials are
// combined in a fixed order by the last block to finish.
#pragma once
#include "common.cuh"

namespace nsb {

constexpr int kRedBlocks = kNumSM * 4;  // grid of every reduction kernel
constexpr int kRedThreads = 256;
constexpr int kMaxDots = 64;            // restart length + 2 at most

template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double *smem /* NV*8 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    if (lane == 0) smem[j * (kRedThreads / 32) + warp] = v[j];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      double s = lane < kRedThreads / 32 ? smem[j * (kRedThreads / 32) + lane] : 0.0;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      v[j] = s;
    }
  }
  __syncthreads();
}

// Active entries of a (possibly distributed) vector: element e of the active
// range lives at e for e < n1 (owned velocity dofs) and at e + gap beyond
// (the replicated pressure part sits behind the velocity ghosts); gap = 0 on a
// single GPU.
__device__ __forceinline__ int64_t active_index(int64_t e, int64_t n1, int64_t gap) { return e < n1 ? e : e + gap; }

// The three passes of one CGS2 orthogonalisation step, all instances of one kernel:
//   MODE 0:  out[i] = V_i . w (i < k), out[k] = w . w when with_self
//   MODE 1:  w += sign * V coef, then out[i] = V_i . w of the UPDATED w -- the basis is read once
//            for the first projection and the second set of inner products
//   MODE 2:  w += sign * V coef, then out[0] = w . w when with_self
// K >= k is the compile-time width (multiple of 4): every thread keeps the K basis entries of an
// element in registers, so K + 1 independent loads per element are in flight (deep memory-level
// parallelism is what an HBM-bound reduction over k+1 streams needs), U elements per thread.
// Updates run over the first n_upd active entries, reductions over the first n_dot (on several
// GPUs the replicated pressure part is updated everywhere but counted on one rank only).
// partials: (K+1) * gridDim.x doubles; counter: one unsigned, zero on entry, reset on exit.
// Reductions are deterministic: block partials are summed in a fixed order by the last block.
// (Requesting the next trip's lines with prefetch.global.L2, which pays in the slab kernels, costs here: both
// passes against 14 vectors 0.453 -> 0.563 ms on B200 -- K + 1 requests per element are too many.  Splitting the
// projection's chain of K dependent DFMAs into four partial sums with the coefficients in registers (the
// source-level samples show 25 % fixed-latency stalls there) costs 26 more registers and is slower too: 0.490 ms.)
template <int K, int U, int MODE>
__global__ void __launch_bounds__(kRedThreads)
    ortho_kernel(const double *__restrict__ V, int64_t ld, int k, const double *__restrict__ coef, double sign,
                 double *w, int64_t n_upd, int64_t n_dot, int64_t n1, int64_t gap, int with_self,
                 double *__restrict__ out, double *partials, unsigned *counter) {
  __shared__ double s_coef[K];
  __shared__ double smem[kRedThreads / 32];
  __shared__ bool is_last;
  if (MODE != 0) {
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_coef[i] = i < k ? sign * coef[i] : 0.0;
    __syncthreads();
  }
  constexpr int NA = MODE == 2 ? 1 : K;
  double acc[NA], self = 0.0;
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = 0.0;
  const int64_t n = n_upd > n_dot ? n_upd : n_dot;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < n; t0 += U * stride) {
    double wv[U], v[U][K];
    int64_t e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t t = t0 + u * stride;
      e[u] = active_index(t < n ? t : n - 1, n1, gap);
      wv[u] = t < n ? w[e[u]] : 0.0;
#pragma unroll
      for (int i = 0; i < K; ++i) v[u][i] = (i < k && t < n) ? __ldcs(V + (int64_t)i * ld + e[u]) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t t = t0 + u * stride;
      double a = wv[u];
      if (MODE != 0) {
#pragma unroll
        for (int i = 0; i < K; ++i) a += s_coef[i] * v[u][i];
        if (t < n_upd) w[e[u]] = a;
      }
      if (t < n_dot) {
        if (MODE != 2) {
#pragma unroll
          for (int i = 0; i < K; ++i) acc[i] += v[u][i] * a;
        }
        self += a * a;
      }
    }
  }
  if (MODE == 2 && !with_self) return;
  const int total = MODE == 2 ? 1 : k + (with_self ? 1 : 0);
  // block partials: acc[0..k) then self
  if (MODE != 2) {
#pragma unroll
    for (int i = 0; i < K; ++i) {
      if (i < k) {  // k is uniform
        double one[1] = {acc[i]};
        block_reduce<1>(one, smem);
        if (threadIdx.x == 0) partials[(int64_t)i * gridDim.x + blockIdx.x] = one[0];
      }
    }
  }
  if (MODE == 2 || with_self) {
    double one[1] = {self};
    block_reduce<1>(one, smem);
    if (threadIdx.x == 0) partials[(int64_t)(MODE == 2 ? 0 : k) * gridDim.x + blockIdx.x] = one[0];
  }
  __threadfence();
  if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int i = threadIdx.x >> 5; i < total; i += kRedThreads / 32) {  // one warp per output, fixed order
      double s = 0;
      for (int b = threadIdx.x & 31; b < (int)gridDim.x; b += 32) s += partials[(int64_t)i * gridDim.x + b];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) == 0) out[i] = s;
    }
    if (threadIdx.x == 0) *counter = 0;
  }
}

// y = a * x  (in place when y == x)
__global__ void scale_kernel(int64_t n, double a, const double *x, double *y) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    y[e] = a * x[e];
}
// y = x / sqrt(*norm2)   (normalisation with a device-resident norm)
__global__ void normalize_kernel(int64_t n, const double *__restrict__ norm2, const double *x, double *y) {
  const double a = rsqrt(*norm2);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    y[e] = a * x[e];
}
// y = a*x + b*y
__global__ void axpby_kernel(int64_t n, double a, const double *__restrict__ x, double b, double *__restrict__ y) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    y[e] = a * x[e] + b * y[e];
}

}  // namespace nsb
