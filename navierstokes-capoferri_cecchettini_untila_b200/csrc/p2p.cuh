// One-sided exchanges between the GPUs of one box over NVLink / NVSwitch peer memory.
//
// What it replaces: the Epetra Import of ghost values before every product of the
// reference (src/NavierStokes.cpp:377, 982, 992; `solution = solution_owned`, :395)
// and the replication of the owned pressure rows.  Round 1 did these with grouped
// ncclSend/ncclRecv and ncclBroadcast calls: ~12 latency-bound NCCL launches per
// outer GMRES iteration, ~35 us each, a third of the iteration at 8 GPUs.
//
// Here every rank owns an ARENA (one cudaMalloc, exported with cudaIpcGetMemHandle
// and mapped by its peers) that holds, per channel, two receive staging buffers and
// one arrival flag per peer.  An exchange is two short kernels on the rank's stream:
//   push:   gather the owned entries a peer needs and STORE them straight into that
//           peer's staging buffer (peer address: NVLink), then one system-scope
//           release store of the exchange number into the peer's flag;
//   unpack: acquire-spin until every peer's flag shows this exchange number, then
//           copy the staging buffer into the ghost slots of the destination vector.
// No host involvement, no NCCL, capturable in a CUDA graph (the exchange number
// lives in device memory and is advanced by the unpack kernel).
//
// Hazards.  Staging is double buffered by the parity of the exchange number e.  A
// sender can only reach push(e+2) after its own unpack(e+1), which needs the
// receiver's push(e+1), which the receiver enqueued after its unpack(e): so nobody
// overwrites staging[e & 1] while its owner still reads exchange e.  This needs
// symmetric neighbour lists and the same sequence of exchanges on all ranks (SPMD),
// both true here.  A spin gives up after kP2PTimeoutNs and raises the context's
// error flag instead of hanging the GPU.
#pragma once
#include "common.cuh"

namespace nsb {

constexpr int kP2PMaxPeers = 16;
constexpr int kP2PChannels = 5;    // see nsb_ctx
constexpr int kP2PReduceSlot = 128;  // doubles per rank in the all-reduce staging (>= 2 * restart + 2)
constexpr unsigned long long kP2PTimeoutNs = 5ull * 1000ull * 1000ull * 1000ull;

struct P2PState {               // device memory, one per channel
  unsigned long long epoch;     // number of completed exchanges
  unsigned int push_count, unpack_count;
  int error;                    // 1: a peer's flag did not arrive in time
};

// kernel arguments of one channel (by value: < 1 KB)
struct P2PArgs {
  int n_peers;
  int peer_rank[kP2PMaxPeers];
  int64_t send_ptr[kP2PMaxPeers + 1];             // entries sent to peer k: [send_ptr[k], send_ptr[k+1])
  double *peer_stage[kP2PMaxPeers];               // remote staging base (parity 0)
  int64_t peer_cap[kP2PMaxPeers];                 // doubles per parity of the remote staging
  int64_t peer_off[kP2PMaxPeers];                 // first ENTRY of my block inside the remote staging
  unsigned long long *peer_flag[kP2PMaxPeers];    // remote flags[my rank]
  const unsigned long long *my_flags;             // local flags, indexed by peer rank
  const double *my_stage;                         // local staging base (parity 0)
  int64_t my_cap;
  P2PState *state;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// push: entry i of the send list (sent to peer k, send_ptr[k] <= i < send_ptr[k+1]) is
// x[width * (idx ? idx[i] : base + (i - send_ptr[k])) + c], c < width  -- idx == nullptr: every peer gets the same
// contiguous block starting at `base` (all-gather)
__global__ void __launch_bounds__(256) p2p_push_kernel(P2PArgs a, int width, const uint32_t *__restrict__ idx,
                                                       int64_t base, const double *__restrict__ x) {
  const unsigned long long e = a.state->epoch + 1;
  const int64_t par = (int64_t)(e & 1ull);
  const int64_t total = a.send_ptr[a.n_peers] * width;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / width;
    const int c = (int)(t - i * width);
    int k = 0;
    while (k + 1 < a.n_peers && i >= a.send_ptr[k + 1]) ++k;
    const int64_t src = idx != nullptr ? (int64_t)idx[i] : base + (i - a.send_ptr[k]);
    a.peer_stage[k][par * a.peer_cap[k] + (a.peer_off[k] + (i - a.send_ptr[k])) * width + c] = x[src * width + c];
  }
  // one system-scope fence per CTA (after the barrier it covers the stores of all its threads; a fence in every
  // thread cost ~5 us per exchange), then the last CTA to arrive releases the flags
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = atomicAdd(&a.state->push_count, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    if ((int)threadIdx.x < a.n_peers) {
      __threadfence_system();
      st_release_sys(a.peer_flag[threadIdx.x], e);
    }
    if (threadIdx.x == 0) a.state->push_count = 0;
  }
}

// unpack: wait for all peers, then entry j of the staging buffer goes to
// y[width * (idx ? idx[j] : base + j) + c]; entries [skip_begin, skip_end) are left alone (the own block of
// an all-gather).  n_entries: entries of the local staging in use.
__global__ void __launch_bounds__(256) p2p_unpack_kernel(P2PArgs a, int width, const uint32_t *__restrict__ idx,
                                                         int64_t base, int64_t n_entries, int64_t skip_begin,
                                                         int64_t skip_end, double *__restrict__ y) {
  const unsigned long long e = a.state->epoch + 1;
  const int64_t par = (int64_t)(e & 1ull);
  if ((int)threadIdx.x < a.n_peers) {
    const unsigned long long *f = a.my_flags + a.peer_rank[threadIdx.x];
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) < e) {
      if (global_ns() - t0 > kP2PTimeoutNs) {
        a.state->error = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  const double *st = a.my_stage + par * a.my_cap;
  const int64_t total = n_entries * width;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = t / width;
    if (j >= skip_begin && j < skip_end) continue;
    const int c = (int)(t - j * width);
    const int64_t dst = idx != nullptr ? (int64_t)idx[j] : base + j;
    y[dst * width + c] = __ldcv(st + t);
  }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(&a.state->unpack_count, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    a.state->unpack_count = 0;
    a.state->epoch = e;
  }
}

// push and unpack of one exchange in ONE launch (every launch costs 2-3 us inside the captured graph and the
// iteration at 8 GPUs holds 12 exchanges): every CTA first stores its share of the send list into the peers'
// staging, the last CTA to finish releases the flags, then every CTA waits for the peers' flags and copies its
// share of the staging into the ghost slots.  A CTA never waits for CTAs of its own grid (only the flag release
// depends on all of them, and nobody of this grid waits for it), so the grid need not be co-resident.
__global__ void __launch_bounds__(256) p2p_exchange_kernel(P2PArgs a, int width, const uint32_t *__restrict__ send_idx,
                                                           int64_t send_base, const double *x,
                                                           const uint32_t *__restrict__ recv_idx, int64_t recv_base,
                                                           int64_t n_entries, int64_t skip_begin, int64_t skip_end,
                                                           double *y) {  // y may be x (owned entries read, ghost slots written)
  const unsigned long long e = a.state->epoch + 1;
  const int64_t par = (int64_t)(e & 1ull);
  __shared__ bool last;
  {
    const int64_t total = a.send_ptr[a.n_peers] * width;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
      const int64_t i = t / width;
      const int c = (int)(t - i * width);
      int k = 0;
      while (k + 1 < a.n_peers && i >= a.send_ptr[k + 1]) ++k;
      const int64_t src = send_idx != nullptr ? (int64_t)send_idx[i] : send_base + (i - a.send_ptr[k]);
      a.peer_stage[k][par * a.peer_cap[k] + (a.peer_off[k] + (i - a.send_ptr[k])) * width + c] = x[src * width + c];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      last = atomicAdd(&a.state->push_count, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
      if ((int)threadIdx.x < a.n_peers) {
        __threadfence_system();
        st_release_sys(a.peer_flag[threadIdx.x], e);
      }
      if (threadIdx.x == 0) a.state->push_count = 0;
    }
  }
  if ((int)threadIdx.x < a.n_peers) {
    const unsigned long long *f = a.my_flags + a.peer_rank[threadIdx.x];
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) < e) {
      if (global_ns() - t0 > kP2PTimeoutNs) {
        a.state->error = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  const double *st = a.my_stage + par * a.my_cap;
  const int64_t total = n_entries * width;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = t / width;
    if (j >= skip_begin && j < skip_end) continue;
    const int c = (int)(t - j * width);
    const int64_t dst = recv_idx != nullptr ? (int64_t)recv_idx[j] : recv_base + j;
    y[dst * width + c] = __ldcv(st + t);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(&a.state->unpack_count, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    a.state->unpack_count = 0;
    a.state->epoch = e;
  }
}

// all-reduce of `count` <= kP2PReduceSlot doubles: every rank has pushed its partial sums into slot `my rank`
// of all peers' staging (p2p_push_kernel, contiguous mode); sum the slots in rank order -- the same order on
// every rank, so all ranks get bit-identical results and take the same branches -- and write buf in place.
__global__ void __launch_bounds__(kP2PReduceSlot) p2p_reduce_kernel(P2PArgs a, int count, int my_rank, int n_ranks,
                                                                   double *__restrict__ buf) {
  const unsigned long long e = a.state->epoch + 1;
  const int64_t par = (int64_t)(e & 1ull);
  if ((int)threadIdx.x < a.n_peers) {
    const unsigned long long *f = a.my_flags + a.peer_rank[threadIdx.x];
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) < e) {
      if (global_ns() - t0 > kP2PTimeoutNs) {
        a.state->error = 1;
        break;
      }
      __nanosleep(32);
    }
  }
  __syncthreads();
  const double *st = a.my_stage + par * a.my_cap;
  const int i = threadIdx.x;
  if (i < count) {
    double s = 0.0;
    for (int r = 0; r < n_ranks; ++r) s += r == my_rank ? buf[i] : __ldcv(st + (int64_t)r * kP2PReduceSlot + i);
    buf[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) a.state->epoch = e;
}

// the same all-reduce in one launch: push the partial sums into slot `my rank` of every peer, release, wait, sum
__global__ void __launch_bounds__(kP2PReduceSlot) p2p_allreduce_kernel(P2PArgs a, int count, int my_rank, int n_ranks,
                                                                      double *__restrict__ buf) {
  const unsigned long long e = a.state->epoch + 1;
  const int64_t par = (int64_t)(e & 1ull);
  for (int t = threadIdx.x; t < a.n_peers * count; t += blockDim.x) {
    const int k = t / count, i = t - k * count;
    a.peer_stage[k][par * a.peer_cap[k] + a.peer_off[k] + i] = buf[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < a.n_peers) {
    __threadfence_system();
    st_release_sys(a.peer_flag[threadIdx.x], e);
    const unsigned long long *f = a.my_flags + a.peer_rank[threadIdx.x];
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) < e) {
      if (global_ns() - t0 > kP2PTimeoutNs) {
        a.state->error = 1;
        break;
      }
      __nanosleep(32);
    }
  }
  __syncthreads();
  const double *st = a.my_stage + par * a.my_cap;
  const int i = threadIdx.x;
  if (i < count) {
    double s = 0.0;
    for (int r = 0; r < n_ranks; ++r) s += r == my_rank ? buf[i] : __ldcv(st + (int64_t)r * kP2PReduceSlot + i);
    buf[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) a.state->epoch = e;
}

// Host-side description of one channel.
struct P2PChannel {
  bool ready = false;
  P2PArgs args{};
  int64_t n_send = 0, n_recv = 0;   // entries
  const uint32_t *send_idx = nullptr, *recv_idx = nullptr;  // device lists (nullptr: contiguous blocks)
  DevBuf<uint32_t> own_send_idx, own_recv_idx;              // storage when the channel owns its lists
};

}  // namespace nsb
