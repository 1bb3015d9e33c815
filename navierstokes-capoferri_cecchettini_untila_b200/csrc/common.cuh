// Shared definitions of the device library: error handling, owned device
// buffers, the CSR view kernels take, and the context object behind nsb_ctx.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nsb.h"
#include "fe_tables.h"

namespace nsb {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct ArgError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct StructError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct NoConvergence : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct NcclError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define NSB_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      throw ::nsb::CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ \
                             ":" + std::to_string(__LINE__) + ")");                             \
  } while (0)

// Owning device array.  `bytes_total` tracks the context's footprint.
template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  int64_t *bytes_total = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    if (p && bytes_total) *bytes_total -= (int64_t)(n * sizeof(T));
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count, int64_t *total = nullptr) {
    release();
    bytes_total = total;
    n = count;
    if (count) NSB_CUDA(cudaMalloc(&p, count * sizeof(T)));
    if (bytes_total) *bytes_total += (int64_t)(count * sizeof(T));
  }
  void upload(const T *h, size_t count, cudaStream_t s, int64_t *total = nullptr) {
    if (count != n || !p) alloc(count, total ? total : bytes_total);
    if (count) NSB_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void download(T *h, cudaStream_t s) const {
    if (n) NSB_CUDA(cudaMemcpyAsync(h, p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
    NSB_CUDA(cudaStreamSynchronize(s));
  }
  void zero(cudaStream_t s) {
    if (n) NSB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
};

// What kernels see of a CSR block.
struct CsrView {
  const int64_t *rowptr;
  const uint32_t *colind;
  double *val;
  int64_t n_rows;
};

struct CsrDev {
  int64_t n_rows = 0, n_cols = 0, nnz = 0;
  DevBuf<int64_t> rowptr;
  DevBuf<uint32_t> colind;
  DevBuf<double> val;
  bool have = false;
  CsrView view() const { return {rowptr.p, colind.p, val.p, n_rows}; }
};

constexpr int kNumSM = 148;  // B200

}  // namespace nsb
