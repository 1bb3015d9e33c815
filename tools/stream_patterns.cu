// Experiment (not part of the product): read bandwidth of the access patterns the slab kernels could use.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/stream_patterns tools/stream_patterns.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// P0: linear, 16 B per thread, grid-stride, unroll 4
__global__ void __launch_bounds__(256) p_linear(const double2 *__restrict__ a, int64_t n2, double *out) {
  double s = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    double2 v0 = __ldcs(a + i), v1 = __ldcs(a + i + stride), v2 = __ldcs(a + i + 2 * stride), v3 = __ldcs(a + i + 3 * stride);
    s += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
  }
  for (; i < n2; i += stride) { double2 v = __ldcs(a + i); s += v.x + v.y; }
  if (s == 1.2345e-300) out[0] = s;
}
// P1/P4: CTA = slab of 8 slices, slice = W steps of 32 x 8 B; warp walks its slice; BATCH loads in flight; optional 2 B stream
template <int BATCH, bool IDX>
__global__ void __launch_bounds__(256) p_slice(const double *__restrict__ a, const uint16_t *__restrict__ ix, int W, double *out) {
  const int64_t base = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * W * 32 + (threadIdx.x & 31);
  double s = 0;
  for (int k = 0; k < W; k += BATCH) {
    double v[BATCH]; unsigned j[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) { v[u] = k + u < W ? __ldcs(a + base + 32 * (k + u)) : 0.0; j[u] = (IDX && k + u < W) ? __ldcs(ix + base + 32 * (k + u)) : 0u; }
#pragma unroll
    for (int u = 0; u < BATCH; ++u) s += v[u] * (double)(j[u] + 1);
  }
  if (s == 1.2345e-300) out[0] = s;
}
// P3: CTA-interleaved: at step k the 8 warps read 8 adjacent 256 B chunks
template <int BATCH, bool IDX>
__global__ void __launch_bounds__(256) p_cta(const double *__restrict__ a, const uint16_t *__restrict__ ix, int W, double *out) {
  const int64_t base = (int64_t)blockIdx.x * 8 * W * 32 + threadIdx.x;
  double s = 0;
  for (int k = 0; k < W; k += BATCH) {
    double v[BATCH]; unsigned j[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) { v[u] = k + u < W ? __ldcs(a + base + 256 * (k + u)) : 0.0; j[u] = (IDX && k + u < W) ? __ldcs(ix + base + 256 * (k + u)) : 0u; }
#pragma unroll
    for (int u = 0; u < BATCH; ++u) s += v[u] * (double)(j[u] + 1);
  }
  if (s == 1.2345e-300) out[0] = s;
}
// P5: persistent variant of P3: gridDim = SMs * k CTAs, each walks many slabs (no CTA launch gaps)
template <int BATCH>
__global__ void __launch_bounds__(256) p_cta_persistent(const double *__restrict__ a, int W, int n_slabs, double *out) {
  double s = 0;
  for (int sl = blockIdx.x; sl < n_slabs; sl += gridDim.x) {
    const int64_t base = (int64_t)sl * 8 * W * 32 + threadIdx.x;
    for (int k = 0; k < W; k += BATCH) {
      double v[BATCH];
#pragma unroll
      for (int u = 0; u < BATCH; ++u) v[u] = k + u < W ? __ldcs(a + base + 256 * (k + u)) : 0.0;
#pragma unroll
      for (int u = 0; u < BATCH; ++u) s += v[u];
    }
  }
  if (s == 1.2345e-300) out[0] = s;
}
// P6: copy (read + write), 16 B per thread
__global__ void __launch_bounds__(256) p_copy(const double2 *__restrict__ a, double2 *__restrict__ b, int64_t n2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) b[i] = a[i];
}

int main() {
  const int W = 22;
  const int n_slabs = 15000 * 4;                       // 4x the sweep's matrix so that launch overheads vanish
  const int64_t n = (int64_t)n_slabs * 8 * W * 32;     // doubles
  double *a, *b, *out; uint16_t *ix;
  CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8)); CK(cudaMalloc(&ix, n * 2)); CK(cudaMalloc(&out, 8));
  CK(cudaMemset(a, 0, n * 8)); CK(cudaMemset(ix, 0, n * 2));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char *name, double bytes, auto launch) {
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.3f ms  %7.1f GB/s  (%s)\n", name, ms / reps, bytes * reps / (ms * 1e6), cudaGetErrorString(cudaGetLastError()));
  };
  const double vb = (double)n * 8, ib = (double)n * 2;
  timeit("P0 linear 16B/thread, 148*8 CTAs", vb, [&] { p_linear<<<148 * 8, 256>>>((const double2 *)a, n / 2, out); });
  timeit("P0 linear 16B/thread, 148*32 CTAs", vb, [&] { p_linear<<<148 * 32, 256>>>((const double2 *)a, n / 2, out); });
  timeit("P6 copy 16B/thread (read+write bytes)", 2 * vb, [&] { p_copy<<<148 * 16, 256>>>((const double2 *)a, (double2 *)b, n / 2); });
  timeit("P1 slice 8B, batch 4", vb, [&] { p_slice<4, false><<<n_slabs, 256>>>(a, ix, W, out); });
  timeit("P1 slice 8B, batch 8", vb, [&] { p_slice<8, false><<<n_slabs, 256>>>(a, ix, W, out); });
  timeit("P1 slice 8B, batch 22", vb, [&] { p_slice<22, false><<<n_slabs, 256>>>(a, ix, W, out); });
  timeit("P4 slice 8B+2B, batch 4", vb + ib, [&] { p_slice<4, true><<<n_slabs, 256>>>(a, ix, W, out); });
  timeit("P4 slice 8B+2B, batch 8", vb + ib, [&] { p_slice<8, true><<<n_slabs, 256>>>(a, ix, W, out); });
  timeit("P3 cta-interleaved 8B, batch 4", vb, [&] { p_cta<4, false><<<n_slabs, 256>>>(a, ix, W, out); });
  timeit("P3 cta-interleaved 8B, batch 8", vb, [&] { p_cta<8, false><<<n_slabs, 256>>>(a, ix, W, out); });
  timeit("P3 cta-interleaved 8B+2B, batch 8", vb + ib, [&] { p_cta<8, true><<<n_slabs, 256>>>(a, ix, W, out); });
  timeit("P5 persistent cta-interleaved, 148*6, b8", vb, [&] { p_cta_persistent<8><<<148 * 6, 256>>>(a, W, n_slabs, out); });
  timeit("P5 persistent cta-interleaved, 148*8, b4", vb, [&] { p_cta_persistent<4><<<148 * 8, 256>>>(a, W, n_slabs, out); });
  return 0;
}
