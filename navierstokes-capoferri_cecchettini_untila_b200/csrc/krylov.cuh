// Fused vector kernels of the GMRES iteration (replaces the l2_norm / dot /
// add / sadd / scale calls deal.II's SolverGMRES and
// PreconditionASIMPLE::vmult issue one at a time, reference
// src/NavierStokes.cpp:348, 377, 978-994).
//
// Orthogonalisation is classical Gram-Schmidt applied twice (CGS2): one
// kernel forms all k inner products V^T w reading w and each basis vector
// once, one kernel applies w -= V h.  All reductions are deterministic: block
// partials are combined in a fixed order by the last block to finish.
#pragma once
#include "common.cuh"

namespace nsb {

constexpr int kDotChunk = 8;
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

// out[i] = V_i . w for i < k (V_i = V + i*ld); out[k] = w . w when with_self.
// partials: kMaxDots * gridDim.x doubles; counter: one unsigned, zero on entry
// and reset to zero on exit.
__global__ void __launch_bounds__(kRedThreads) multi_dot_kernel(const double *__restrict__ V, int64_t ld, int k,
                                                                const double *__restrict__ w, int64_t n, int64_t n1,
                                                                int64_t gap, int with_self,
                                                                double *__restrict__ out, double *partials,
                                                                unsigned *counter) {
  __shared__ double smem[kDotChunk * (kRedThreads / 32)];
  __shared__ bool is_last;
  const int total = k + (with_self ? 1 : 0);
  for (int c0 = 0; c0 < total; c0 += kDotChunk) {
    double acc[kDotChunk];
#pragma unroll
    for (int j = 0; j < kDotChunk; ++j) acc[j] = 0.0;
    const int cnt = min(kDotChunk, total - c0);
    for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += (int64_t)gridDim.x * blockDim.x) {
      const int64_t e = active_index(a, n1, gap);
      const double we = w[e];
#pragma unroll
      for (int j = 0; j < kDotChunk; ++j)
        if (j < cnt) {
          const int i = c0 + j;
          acc[j] += (i < k ? V[(int64_t)i * ld + e] : we) * we;
        }
    }
    block_reduce<kDotChunk>(acc, smem);
    if (threadIdx.x == 0)
#pragma unroll
      for (int j = 0; j < kDotChunk; ++j)
        if (j < cnt) partials[(int64_t)(c0 + j) * gridDim.x + blockIdx.x] = acc[j];
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

// w += sum_{i<k} sign * coef[i] * V_i ;  optionally out_norm2 = w . w afterwards
__global__ void __launch_bounds__(kRedThreads) multi_axpy_kernel(const double *__restrict__ V, int64_t ld, int k,
                                                                 const double *__restrict__ coef, double sign,
                                                                 double *__restrict__ w, int64_t n, int64_t n1,
                                                                 int64_t gap, int64_t n_norm, int with_norm,
                                                                 double *__restrict__ out_norm2, double *partials,
                                                                 unsigned *counter) {
  __shared__ double s_coef[kMaxDots];
  __shared__ double smem[kRedThreads / 32];
  __shared__ bool is_last;
  for (int i = threadIdx.x; i < k; i += blockDim.x) s_coef[i] = sign * coef[i];
  __syncthreads();
  double nrm[1] = {0.0};
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = active_index(t, n1, gap);
    double a = w[e];
#pragma unroll 4
    for (int i = 0; i < k; ++i) a += s_coef[i] * V[(int64_t)i * ld + e];
    w[e] = a;
    if (t < n_norm) nrm[0] += a * a;  // the replicated part is counted on one rank only
  }
  if (!with_norm) return;
  block_reduce<1>(nrm, smem);
  if (threadIdx.x == 0) partials[blockIdx.x] = nrm[0];
  __threadfence();
  if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    double s = 0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) s += partials[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
      *out_norm2 = s;
      *counter = 0;
    }
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
