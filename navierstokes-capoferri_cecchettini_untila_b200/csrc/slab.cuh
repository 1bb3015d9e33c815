// Slab storage of the node-level velocity block F_s (A00 = F_s (x) I_dim) and the
// kernels that stream it: the velocity rows of y = A x (GMRES, reference
// src/NavierStokes.cpp:377) and the Jacobi-type inner sweep on F that replaces
// the ILU-preconditioned inner GMRES of PreconditionASIMPLE::vmult (:978-981).
//
// Why not CSR.  With 12 B per stored non-zero for 2*dim flops the node-block CSR
// kernels were bound by the GATHER, not by the matrix stream: every non-zero
// pulled a 32-byte sector of the vector through L1/L2 (ncu, round 1: L1/LSU the
// busiest unit, 91 % long-scoreboard stalls, 34 % of the HBM roofline), and rows
// of 18 / 27 / 65 entries (edge / face-adjacent / vertex nodes) left most lanes
// of a sub-warp-per-row kernel idle.
//
// Layout.  Consecutive rows are grouped into slabs of at most kSlabThreads
// *virtual rows* (a row longer than ~1.4x the mean is cut into <= 3 chunks of
// equal length so that all virtual rows have similar lengths).  Per slab:
//   * a WINDOW: the sorted list of distinct column nodes of its rows
//     (~3-4 per row).  The kernel stages the dim values of every window node in
//     shared memory once, with coalesced loads, so every vector sector is
//     fetched once per slab instead of once per non-zero;
//   * the virtual rows sorted by length and stored as 8 ELL slices of 32 rows
//     (entry-major inside a slice, two entries of a lane adjacent): thread t of
//     the CTA owns virtual row t and streams `val` (8 B) and a 16-bit
//     window-local column index (2 B) with coalesced 512 B + 128 B warp loads
//     -- 10 B per non-zero instead of 12 B, no row pointers, no shuffles;
//   * the order of the entries inside a virtual row is chosen so that the 16
//     lanes of a half-warp read different shared-memory bank pairs (pass 3 of
//     build_slabs);
//   * per real row a packed word with the thread positions of its <= 3 chunks;
//     partial sums are combined in a fixed order (deterministic) in the
//     coalesced epilogue that also applies the vector update of the sweep.
// The CSR values stay the assembly target (scatter by slot); `slab_repack_kernel`
// copies them into the ELL order once per time step through a precomputed map.
#pragma once
#include <algorithm>
#include <numeric>

#include "common.cuh"

namespace nsb {

#ifndef NSB_SLAB_THREADS
#define NSB_SLAB_THREADS 256
#endif
constexpr int kSlabThreads = NSB_SLAB_THREADS;  // virtual rows (= threads) per slab
constexpr int kSlabSlices = kSlabThreads / 32;
constexpr int kSlabMaxChunks = 3;
constexpr uint32_t kVposNone = 0x3ffu;
constexpr uint32_t kSlabPad = 0xffffffffu;

struct SlabView {
  int n_slabs;
  const uint32_t *slab_row;  // n_slabs+1: first real row of every slab
  const uint32_t *win_ptr;   // n_slabs+1: offsets into win_list
  const uint32_t *win_list;  // window nodes, ascending inside a slab
  const int64_t *slice_ptr;  // n_slabs*kSlabSlices+1: offsets into val/idx (multiples of 32)
  const double *val;
  const uint16_t *idx;
  const uint32_t *vpos;      // per real row: 3 x 10-bit thread positions of its chunks (kVposNone = unused)
};

// Position of entry k of lane `lane` inside a slice: two consecutive entries of a lane are adjacent, so
// a lane fetches 2 values with one 128-bit load and 2 indices with one 32-bit load (512 B + 128 B per
// warp request; on B200 this measured 1 % faster than 256 B + 64 B requests).  Slice widths are even, which
// also gives the bank-aware schedule slack to avoid conflicts.
__host__ __device__ inline int64_t slab_entry_pos(int k, int lane) { return (int64_t)(k >> 1) * 64 + lane * 2 + (k & 1); }

struct SlabHost {
  std::vector<uint32_t> slab_row, win_ptr, win_list, vpos, src;
  std::vector<int64_t> slice_ptr;
  std::vector<uint16_t> idx;
  uint32_t max_window = 0;
  int64_t nnz = 0;
  int64_t max_slab_entries = 0;  // largest number of stored entries (incl. padding) of one slab
  double bank_wavefronts_per_step = 1.0;  // shared-memory wavefronts per half-warp load (1 = conflict-free)
  double bank_wavefronts_bound = 1.0;     // lower bound for this window order (most loaded residue per half-warp)
};

struct SlabDev {
  int n_slabs = 0;
  int64_t n_rows = 0, nnz = 0, padded = 0, max_slab_entries = 0;
  uint32_t max_window = 0;
  DevBuf<uint32_t> slab_row, win_ptr, win_list, vpos, src;
  DevBuf<int64_t> slice_ptr;
  DevBuf<double> val;
  DevBuf<uint16_t> idx;
  bool have = false;
  SlabView view() const {
    return {n_slabs, slab_row.p, win_ptr.p, win_list.p, slice_ptr.p, val.p, idx.p, vpos.p};
  }
  size_t window_total() const { return win_list.n; }
};

// Host construction from the CSR pattern (once per setup).  window_cap: largest
// window (in nodes) a slab may have -- it bounds the shared memory of the kernels.
inline SlabHost build_slabs(int64_t n_rows, int64_t n_cols, const int64_t *rp, const uint32_t *ci, uint32_t window_cap) {
  SlabHost H;
  H.nnz = rp[n_rows];
  if (H.nnz >= (int64_t)kSlabPad) throw ArgError("slab storage: more than 2^32-2 stored entries per rank");
  const double mean = n_rows ? (double)H.nnz / (double)n_rows : 1.0;
  const int target = std::max(8, (int)std::lround(0.85 * mean));
  auto n_chunks = [&](int64_t len) {
    return (int)std::min<int64_t>(kSlabMaxChunks, std::max<int64_t>(1, (len + target / 2) / target));
  };
  // pass 1 (sequential): slab boundaries and unsorted windows
  std::vector<int32_t> stamp((size_t)n_cols, -1);
  H.slab_row.push_back(0);
  H.win_ptr.push_back(0);
  H.win_list.reserve((size_t)n_rows * 4);
  int32_t slab = 0;
  int vcount = 0;
  uint32_t wcount = 0;
  for (int64_t r = 0; r < n_rows; ++r) {
    const int64_t b = rp[r], e = rp[r + 1];
    if ((uint64_t)(e - b) > window_cap) throw StructError("slab storage: a row has more entries than the window capacity");
    const int nch = n_chunks(e - b);
    uint32_t fresh = 0;
    for (int64_t k = b; k < e; ++k) fresh += stamp[ci[k]] != slab;
    if (vcount > 0 && (vcount + nch > kSlabThreads || wcount + fresh > window_cap)) {
      H.slab_row.push_back((uint32_t)r);
      H.win_ptr.push_back((uint32_t)H.win_list.size());
      ++slab;
      vcount = 0;
      wcount = 0;
    }
    for (int64_t k = b; k < e; ++k)
      if (stamp[ci[k]] != slab) {
        stamp[ci[k]] = slab;
        H.win_list.push_back(ci[k]);
        ++wcount;
      }
    vcount += nch;
  }
  H.slab_row.push_back((uint32_t)n_rows);
  H.win_ptr.push_back((uint32_t)H.win_list.size());
  const int64_t ns = (int64_t)H.slab_row.size() - 1;
  // pass 2 (parallel): sort windows, sort virtual rows by length, slice widths
  struct VRow {
    uint32_t row;
    uint16_t off, len;  // chunk = entries [rp[row]+off, +len)
    uint16_t chunk;     // position among the chunks of the row
  };
  std::vector<std::vector<VRow>> vrows((size_t)ns);
  std::vector<int64_t> slice_len((size_t)ns * kSlabSlices, 0);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t s = 0; s < ns; ++s) {
    std::sort(H.win_list.begin() + H.win_ptr[s], H.win_list.begin() + H.win_ptr[s + 1]);
    std::vector<VRow> &v = vrows[s];
    for (uint32_t r = H.slab_row[s]; r < H.slab_row[s + 1]; ++r) {
      const int64_t len = rp[r + 1] - rp[r];
      const int nch = n_chunks(len);
      int64_t off = 0;
      for (int j = 0; j < nch; ++j) {
        const int64_t l = (len - off + (nch - j) - 1) / (nch - j);
        v.push_back({r, (uint16_t)off, (uint16_t)l, (uint16_t)j});
        off += l;
      }
    }
    std::stable_sort(v.begin(), v.end(), [](const VRow &a, const VRow &b) { return a.len > b.len; });
    for (int w = 0; w < kSlabSlices; ++w)
      slice_len[s * kSlabSlices + w] = (size_t)w * 32 < v.size() ? 32 * (int64_t)((v[(size_t)w * 32].len + 1) & ~1) : 0;
  }
  H.slice_ptr.assign((size_t)ns * kSlabSlices + 1, 0);
  for (size_t i = 0; i < slice_len.size(); ++i) H.slice_ptr[i + 1] = H.slice_ptr[i] + slice_len[i];
  const int64_t total = H.slice_ptr.back();
  for (int64_t sl = 0; sl < ns; ++sl)
    H.max_slab_entries = std::max(H.max_slab_entries, H.slice_ptr[(sl + 1) * kSlabSlices] - H.slice_ptr[sl * kSlabSlices]);
  H.idx.assign((size_t)total, 0);
  H.src.assign((size_t)total, kSlabPad);
  H.vpos.assign((size_t)n_rows, kVposNone | (kVposNone << 10) | (kVposNone << 20));
  // pass 3 (parallel): fill.  The order of the entries inside a virtual row is free, so it is chosen
  // per half-warp (the unit in which a 64-bit shared-memory load is served) such that at every step
  // the 16 lanes read window nodes with distinct (index mod 16), i.e. distinct bank pairs for each of
  // the dim components: a lane takes a free residue when it has one, idles (zero padding) when it has
  // slack in its slice, and only otherwise accepts a bank conflict.  ncu before this ordering: 2.2
  // conflict wavefronts per shared load, L1/shared pipe the busiest unit of the sweep kernel at 73 %.
  uint32_t maxw = 0;
  int64_t wavefronts = 0, steps = 0, bound = 0;
  int sched_error = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(max : maxw, sched_error) reduction(+ : wavefronts, steps, bound)
  for (int64_t s = 0; s < ns; ++s) {
    const uint32_t *wb = H.win_list.data() + H.win_ptr[s], *we = H.win_list.data() + H.win_ptr[s + 1];
    const uint32_t nw = (uint32_t)(we - wb);
    maxw = std::max(maxw, nw);
    const std::vector<VRow> &v = vrows[s];
    for (size_t t = 0; t < v.size(); ++t) {
      uint32_t &vp = H.vpos[v[t].row];  // the chunks of a row live in one slab: no race across threads
      vp = (vp & ~(0x3ffu << (10 * v[t].chunk))) | ((uint32_t)t << (10 * v[t].chunk));
    }
    struct Ent {
      uint16_t idx;
      uint32_t src;
    };
    std::vector<Ent> ent[16];
    for (size_t h0 = 0; h0 < v.size(); h0 += 16) {  // one half-warp
      const int64_t sl = s * kSlabSlices + (int64_t)(h0 >> 5);
      const int W = (int)((H.slice_ptr[sl + 1] - H.slice_ptr[sl]) >> 5);
      const int nl = (int)std::min<size_t>(16, v.size() - h0);
      int cnt[16][16], rem[16], load[16];
      for (int r = 0; r < 16; ++r) load[r] = 0;
      for (int l = 0; l < 16; ++l) {
        rem[l] = 0;
        for (int r = 0; r < 16; ++r) cnt[l][r] = 0;
        ent[l].clear();
        if (l >= nl) continue;
        const VRow &q = v[h0 + l];
        for (int k = 0; k < q.len; ++k) {
          const int64_t p = rp[q.row] + q.off + k;
          ent[l].push_back({(uint16_t)(std::lower_bound(wb, we, ci[p]) - wb), (uint32_t)p});
        }
        // group by residue; inside a residue keep the column order
        std::stable_sort(ent[l].begin(), ent[l].end(), [](const Ent &a, const Ent &b) { return (a.idx & 15) < (b.idx & 15); });
        for (const Ent &e : ent[l]) {
          ++cnt[l][e.idx & 15];
          ++load[e.idx & 15];
        }
        rem[l] = q.len;
      }
      {
        int mxl = W;
        for (int r = 0; r < 16; ++r) mxl = std::max(mxl, load[r]);
        bound += mxl;
      }
      int first[16][16];  // first unused entry of every (lane, residue) group
      for (int l = 0; l < 16; ++l) {
        int o = 0;
        for (int r = 0; r < 16; ++r) {
          first[l][r] = o;
          o += cnt[l][r];
        }
      }
      for (int k = 0; k < W; ++k) {
        int mult[16], order[16], pick[16];
        for (int r = 0; r < 16; ++r) mult[r] = 0;
        for (int l = 0; l < 16; ++l) {
          order[l] = l;
          pick[l] = -1;
        }
        // lanes without slack first, then by remaining length
        std::stable_sort(order, order + 16, [&](int a, int b) { return rem[a] > rem[b]; });
        for (int oi = 0; oi < 16; ++oi) {
          const int l = order[oi];
          if (rem[l] == 0) continue;
          int best = -1;
          for (int r = 0; r < 16; ++r)
            if (cnt[l][r] > 0 && mult[r] == 0 && (best < 0 || load[r] > load[best])) best = r;
          if (best < 0) {
            if (rem[l] < W - k) continue;  // slack: idle this step
            for (int r = 0; r < 16; ++r)
              if (cnt[l][r] > 0 && (best < 0 || mult[r] < mult[best] || (mult[r] == mult[best] && load[r] > load[best])))
                best = r;
          }
          pick[l] = best;
          ++mult[best];
        }
        for (int l = 0; l < 16; ++l) {
          const int64_t p = H.slice_ptr[sl] + slab_entry_pos(k, (int)((h0 + l) & 31));
          if (pick[l] >= 0) {
            const int r = pick[l];
            const Ent &e = ent[l][first[l][r]++];
            --cnt[l][r];
            --load[r];
            --rem[l];
            H.idx[(size_t)p] = e.idx;
            H.src[(size_t)p] = e.src;
          } else {  // padding (value 0): point it at a bank pair nobody uses in this step
            int r = 0;
            while (r < 16 && mult[r] != 0) ++r;
            if (r == 16 || (uint32_t)r >= nw) r = 0;
            ++mult[r];
            H.idx[(size_t)p] = (uint16_t)r;
          }
        }
        int mx = 1;
        for (int r = 0; r < 16; ++r) mx = std::max(mx, mult[r]);
        wavefronts += mx;
        ++steps;
      }
      for (int l = 0; l < 16; ++l)
        if (rem[l] != 0) sched_error = 1;  // (exceptions must not leave an OpenMP region)
    }
  }
  if (sched_error) throw StructError("slab storage: internal scheduling error");
  H.bank_wavefronts_per_step = steps ? (double)wavefronts / (double)steps : 1.0;
  H.bank_wavefronts_bound = steps ? (double)bound / (double)steps : 1.0;
  H.max_window = maxw;
  return H;
}

inline void upload_slabs(const SlabHost &H, int64_t n_rows, SlabDev &D, cudaStream_t s, int64_t *bytes) {
  D.n_slabs = (int)H.slab_row.size() - 1;
  D.n_rows = n_rows;
  D.nnz = H.nnz;
  D.padded = H.slice_ptr.back();
  D.max_window = H.max_window;
  D.max_slab_entries = H.max_slab_entries;
  D.slab_row.upload(H.slab_row.data(), H.slab_row.size(), s, bytes);
  D.win_ptr.upload(H.win_ptr.data(), H.win_ptr.size(), s, bytes);
  D.win_list.upload(H.win_list.data(), H.win_list.size(), s, bytes);
  D.vpos.upload(H.vpos.data(), H.vpos.size(), s, bytes);
  D.src.upload(H.src.data(), H.src.size(), s, bytes);
  D.slice_ptr.upload(H.slice_ptr.data(), H.slice_ptr.size(), s, bytes);
  D.idx.upload(H.idx.data(), H.idx.size(), s, bytes);
  D.val.alloc((size_t)D.padded, bytes);
  D.val.zero(s);
  NSB_CUDA(cudaStreamSynchronize(s));
  D.have = true;
}

// ELL values from the CSR values (after assembly and boundary rows), padding = 0
__global__ void slab_repack_kernel(int64_t n, const uint32_t *__restrict__ src, const double *__restrict__ csr_val,
                                   double *__restrict__ ell_val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint32_t p = __ldcs(src + i);
    ell_val[i] = p != kSlabPad ? csr_val[p] : 0.0;
  }
}

// Stage the window of slab s in shared memory and form this thread's partial row sum.
// Entries are taken kSlabBatch at a time: all (value, index) loads of a batch are issued before the
// first use.  Measured on B200 at 9.7 M DoFs (round 1, variant builds of this file): batch 4 with 6 resident
// CTAs per SM (40 registers) is as fast as batches of 8 or 16 and as 7-8 resident CTAs -- 0.34 ms per sweep in all
// cases; bank-aware entry order: 0.40 -> 0.345 ms.  The -DNSB_SLAB_* macros below are the switches of those A/B
// builds (tools/time_kernels.py with NSB_LIBNSB=<variant>); their defaults are the measured best.
#ifndef NSB_SLAB_BATCH
#define NSB_SLAB_BATCH 4
#endif
constexpr int kSlabBatch = NSB_SLAB_BATCH;
#ifndef NSB_SLAB_MINBLOCKS
#define NSB_SLAB_MINBLOCKS (1536 / NSB_SLAB_THREADS)
#endif
constexpr int kSlabMinBlocks = NSB_SLAB_MINBLOCKS;  // 1536 resident threads per SM (40 registers)
// Software prefetch into L2 (prefetch.global.L2: no register, no scoreboard slot): while batch k is consumed the
// lines of batch k + NSB_SLAB_PF are requested, and the first NSB_SLAB_PF batches plus the vectors of the epilogue
// before the window is staged, so that the serial phases of a CTA (window fill -> stream -> epilogue, which add up
// because at 6 CTAs per SM nothing else hides them) wait on L2 instead of HBM.  0 = off.  Measured on B200 at
// 9.7 M DoFs (round 2, tools/time_kernels.py): sweep 0.316 ms without, 0.259 / 0.263 / 0.272 / 0.284 ms with a
// distance of 1 / 2 / 4 / 8 batches (0.271 without the epilogue lines); batches of 8 or 6 entries with distance 1:
// 0.266 / 0.262; batches of 2 with distance 2: 0.272.  Also tried: the lines of x behind the window of the slab that
// will run in this CTA's place one wave later: 0.262, no gain; 8 or 7 resident CTAs per SM (32 registers, window cap
// 1152 / 1280 nodes) on top of the two-round-trip window fill: 0.2475 / 0.2518 against 0.2460 with 6.
#ifndef NSB_SLAB_PF
#define NSB_SLAB_PF 1
#endif
constexpr int kSlabPrefetch = NSB_SLAB_PF;
#ifndef NSB_SLAB_PF_EPI  // also request the epilogue's operands (0.271 -> 0.259 ms)
#define NSB_SLAB_PF_EPI 1
#endif
#ifndef NSB_SLAB_FILL  // 1: window fill in two round trips (indices, then cp.async copies); 0: element-strided loop
#define NSB_SLAB_FILL 1
#endif
#ifndef NSB_SLAB_WINCAP3  // window capacity of a 3D slab in nodes (24 B each in shared memory)
#define NSB_SLAB_WINCAP3 1408u
#endif
// Look-ahead (in slabs) of the metadata / window-list prefetch below: half a wave of resident CTAs (6 per SM).
// B200, 9.7 M DoFs: sweep 0.2460 ms without, 0.2453 with a full wave (888), 0.2424 with half a wave (444).
#ifndef NSB_SLAB_AHEAD
#define NSB_SLAB_AHEAD (3 * 148)
#endif
// largest window (in nodes) a slab may have: bounds the shared memory of the kernels and the fill's unroll
template <int DIM>
constexpr uint32_t kSlabWindowCap = (DIM == 3 ? NSB_SLAB_WINCAP3 : 2112u) * (kSlabThreads > 256 ? kSlabThreads / 256 : 1);
#ifndef NSB_G_PF  // prefetch distance (batches) of the A01 slabs: g_slab_apply 0.227 -> 0.194 ms with 1, same with 2
#define NSB_G_PF 1
#endif
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// lines of entry pairs [p0, p0 + np) of a warp's slice: 4 lines of values and 1 line of indices per pair
__device__ __forceinline__ void slab_prefetch_pairs(const double2 *vbase, const uint32_t *ibase, int p0, int np, int W2,
                                                    int lane) {
  for (int q = lane; q < 5 * np; q += 32) {
    const int p = p0 + q / 5, l = q % 5;
    if (p < W2) {
      if (l < 4)
        prefetch_l2(vbase + 32 * p + 8 * l);
      else
        prefetch_l2(ibase + 32 * p);
    }
  }
}
template <int DIM, int BATCH>
__device__ __forceinline__ void slab_product(const SlabView &S, int s, const double *__restrict__ x, double *sm,
                                             double (&acc)[DIM]) {
  static_assert(BATCH % 2 == 0, "entries are fetched in pairs");
  const int t = threadIdx.x;
  const int64_t sl = (int64_t)s * kSlabSlices + (t >> 5);
  const int64_t base = S.slice_ptr[sl];
  const int W2 = (int)((S.slice_ptr[sl + 1] - base) >> 6);  // pairs of entries per lane; warp-uniform
  const double2 *__restrict__ v = reinterpret_cast<const double2 *>(S.val + base) + (t & 31);
  const uint32_t *__restrict__ ix = reinterpret_cast<const uint32_t *>(S.idx + base) + (t & 31);
  const uint32_t w0 = S.win_ptr[s], nw = S.win_ptr[s + 1] - w0;
  if (kSlabPrefetch > 0) slab_prefetch_pairs(v - (t & 31), ix - (t & 31), 0, kSlabPrefetch * (BATCH / 2), W2, t & 31);
#if NSB_SLAB_AHEAD
  {  // lines of the metadata and of the window list of the slab that runs NSB_SLAB_AHEAD slabs later
    const int sf = s + NSB_SLAB_AHEAD;
    if (sf < S.n_slabs) {
      const uint32_t wf = S.win_ptr[sf], nwf = S.win_ptr[sf + 1] - wf;
      if (t == 0) {
        prefetch_l2(S.slab_row + sf + 2 * NSB_SLAB_AHEAD);
        prefetch_l2(S.win_ptr + sf + 2 * NSB_SLAB_AHEAD);
        prefetch_l2(S.slice_ptr + (int64_t)(sf + NSB_SLAB_AHEAD) * kSlabSlices);
      }
      if (32u * t < nwf) prefetch_l2(S.win_list + wf + 32 * t);
    }
  }
#endif
#if NSB_SLAB_FILL
  // Window fill in TWO round trips for the whole CTA: every thread first loads the indices of all its window nodes
  // (thread t owns nodes t, t + T, ...), then issues the copies of their DIM values straight into shared memory
  // (cp.async, 8 bytes each: no registers, nothing waits until the group is complete).  The element-strided loop it
  // replaces went through one dependent (index -> value) pair of round trips per unrolled group of four elements;
  // the source-level stall samples of round 2 put 35 % of the warp time before the first barrier.
  {
    constexpr int NPT = (kSlabWindowCap<DIM> + kSlabThreads - 1) / kSlabThreads;
    uint32_t node[NPT];
#pragma unroll
    for (int q = 0; q < NPT; ++q) node[q] = t + q * kSlabThreads < nw ? __ldg(S.win_list + w0 + t + q * kSlabThreads) : 0xffffffffu;
#pragma unroll
    for (int q = 0; q < NPT; ++q)
      if (node[q] != 0xffffffffu) {
        const double *src = x + (size_t)DIM * node[q];
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm + DIM * (t + q * kSlabThreads));
#pragma unroll
        for (int c = 0; c < DIM; ++c)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8 * c), "l"(src + c) : "memory");
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
#pragma unroll
  for (int c = 0; c < DIM; ++c) acc[c] = 0.0;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
#else
  for (uint32_t i = t; i < DIM * nw; i += kSlabThreads) {
    const uint32_t node = __ldg(S.win_list + w0 + i / DIM);
    sm[i] = __ldg(x + (size_t)DIM * node + i % DIM);
  }
#pragma unroll
  for (int c = 0; c < DIM; ++c) acc[c] = 0.0;
  __syncthreads();
#endif
  for (int k = 0; k < W2; k += BATCH / 2) {
    double2 a[BATCH / 2];
    uint32_t j[BATCH / 2];
    if (kSlabPrefetch > 0)
      slab_prefetch_pairs(v - (t & 31), ix - (t & 31), k + kSlabPrefetch * (BATCH / 2), BATCH / 2, W2, t & 31);
#pragma unroll
    for (int u = 0; u < BATCH / 2; ++u) {
      a[u] = k + u < W2 ? __ldcs(v + 32 * (k + u)) : make_double2(0.0, 0.0);
      j[u] = k + u < W2 ? __ldcs(ix + 32 * (k + u)) : 0u;
    }
#pragma unroll
    for (int u = 0; u < BATCH / 2; ++u)
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        acc[c] += a[u].x * sm[DIM * (j[u] & 0xffffu) + c];
        acc[c] += a[u].y * sm[DIM * (j[u] >> 16) + c];
      }
  }
  __syncthreads();  // every warp is done with the window: its space now takes the partial sums
#pragma unroll
  for (int c = 0; c < DIM; ++c) sm[DIM * t + c] = acc[c];
  __syncthreads();
}

// ---------------------------------------------------------------------------
// Bulk-copy (TMA) variant of slab_product.  The stored entries of a slab are one contiguous range of `val`
// and of `idx`, and slice w belongs to warp w alone: lane 0 of every warp arms a warp-private mbarrier and issues
// two cp.async.bulk copies (its slice of val and of idx, 5-9 KB) BEFORE the window is staged, so the matrix stream
// of the CTA is in flight during the window fill and costs neither registers nor scoreboard slots; the warp then
// waits on its barrier and takes values and indices from shared memory.
// MEASURED NEGATIVE (B200, 9.7 M DoFs, round 2): 0.52 ms per sweep with 256-thread slabs (100 KB of shared memory,
// 2 CTAs = 16 warps per SM), 0.48 ms with 128-thread slabs (4 CTAs per SM), against 0.316 ms for the register-staged
// stream and 0.259 ms with the L2 prefetch above: with the whole slab staged the CTA needs ~100 KB, so only two CTAs
// are resident and the latencies of the window gather and of the epilogue are exposed -- the kernel needs many
// resident warps more than it needs the copy engine.  Kept behind NSB_SWEEP_TMA=1 (parity-tested) as the starting
// point of a persistent, double-buffered version; not used by default.  Shared memory of a CTA:
//   [0,128) mbarriers | window / partial sums (win_doubles) | val (max_slab_entries doubles) | idx (uint16)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}

constexpr int kSlabTmaHeader = 128;  // bytes reserved for the mbarriers

template <int DIM, int BATCH>
__device__ __forceinline__ void slab_product_tma(const SlabView &S, int s, const double *__restrict__ x,
                                                 unsigned char *smraw, uint32_t win_doubles, int64_t max_entries,
                                                 double (&acc)[DIM]) {
  static_assert(BATCH % 2 == 0, "entries are fetched in pairs");
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smraw);
  double *sm = reinterpret_cast<double *>(smraw + kSlabTmaHeader);
  double *sval = sm + ((win_doubles + 15u) & ~15u);  // bulk copies need 16-byte aligned destinations
  uint16_t *sidx = reinterpret_cast<uint16_t *>(sval + max_entries);
  const int64_t base0 = S.slice_ptr[(int64_t)s * kSlabSlices];
  const int64_t base = S.slice_ptr[(int64_t)s * kSlabSlices + warp];
  const int n_ent = (int)(S.slice_ptr[(int64_t)s * kSlabSlices + warp + 1] - base);  // multiple of 64
  const int W2 = n_ent >> 6;
  const int off = (int)(base - base0);
  if (lane == 0) {
    mbar_init(bars + warp, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (n_ent > 0) {
      mbar_expect_tx(bars + warp, (unsigned)n_ent * 10u);
      bulk_g2s(sval + off, S.val + base, (unsigned)n_ent * 8u, bars + warp);
      bulk_g2s(sidx + off, S.idx + base, (unsigned)n_ent * 2u, bars + warp);
    }
  }
  const uint32_t w0 = S.win_ptr[s], nw = S.win_ptr[s + 1] - w0;
  for (uint32_t i = t; i < DIM * nw; i += kSlabThreads) {
    const uint32_t node = __ldg(S.win_list + w0 + i / DIM);
    sm[i] = __ldg(x + (size_t)DIM * node + i % DIM);
  }
#pragma unroll
  for (int c = 0; c < DIM; ++c) acc[c] = 0.0;
  __syncthreads();  // window staged; also orders lane 0's mbarrier init before the other lanes' wait
  if (n_ent > 0) mbar_wait(bars + warp, 0);
  const double2 *v = reinterpret_cast<const double2 *>(sval + off) + lane;
  const uint32_t *ix = reinterpret_cast<const uint32_t *>(sidx + off) + lane;
  for (int k = 0; k < W2; k += BATCH / 2) {
    double2 a[BATCH / 2];
    uint32_t j[BATCH / 2];
#pragma unroll
    for (int u = 0; u < BATCH / 2; ++u) {
      a[u] = k + u < W2 ? v[32 * (k + u)] : make_double2(0.0, 0.0);
      j[u] = k + u < W2 ? ix[32 * (k + u)] : 0u;
    }
#pragma unroll
    for (int u = 0; u < BATCH / 2; ++u)
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        acc[c] += a[u].x * sm[DIM * (j[u] & 0xffffu) + c];
        acc[c] += a[u].y * sm[DIM * (j[u] >> 16) + c];
      }
  }
  __syncthreads();  // every warp is done with the window: its space now takes the partial sums
#pragma unroll
  for (int c = 0; c < DIM; ++c) sm[DIM * t + c] = acc[c];
  __syncthreads();
}

// (F x)[row r of the slab][component c] from the partial sums in shared memory, fixed order
template <int DIM>
__device__ __forceinline__ double slab_row_sum(const double *sm, uint32_t vp, int c) {
  double s = sm[DIM * (vp & 0x3ffu) + c];
  const uint32_t p1 = (vp >> 10) & 0x3ffu, p2 = (vp >> 20) & 0x3ffu;
  if (p1 != kVposNone) s += sm[DIM * p1 + c];
  if (p2 != kVposNone) s += sm[DIM * p2 + c];
  return s;
}

// One Chebyshev-Jacobi sweep on F z = b in three-term form (d_k = z_k - z_{k-1} is not stored):
//   znew = z + c1 * (z - zold) + c2 * (bd - Dinv .* (F z)),   bd = Dinv .* b (formed once per solve)
// Dinv is kept per NODE (the diagonal of A00 = F_s (x) I_dim is the same for the dim components), so
// the vector traffic of a sweep is z, zold, bd read and znew written: 4 1/3 streams instead of the 6
// (b, Dinv, d read; d, znew written; z) of the two-term form.  z, zold, znew are distinct buffers.
template <int DIM>
__global__ void __launch_bounds__(kSlabThreads, kSlabMinBlocks)
    fs_slab_sweep_kernel(SlabView S, const double *__restrict__ dinv_node, const double *__restrict__ bd,
                         const double *__restrict__ z, const double *__restrict__ zold, double *__restrict__ znew,
                         double c1, double c2) {
  extern __shared__ double sm[];
  const int s = blockIdx.x;
  double acc[DIM];
  const uint32_t r0 = S.slab_row[s], nr = S.slab_row[s + 1] - r0;
  if (kSlabPrefetch > 0 && NSB_SLAB_PF_EPI) {  // the epilogue's operands: 16 doubles per line
    for (uint32_t i = 16 * threadIdx.x; i < DIM * nr; i += 16 * kSlabThreads) {
      prefetch_l2(z + (int64_t)DIM * r0 + i);
      if (zold != nullptr) prefetch_l2(zold + (int64_t)DIM * r0 + i);
      prefetch_l2(bd + (int64_t)DIM * r0 + i);
    }
    if (threadIdx.x < (nr + 15) / 16) prefetch_l2(dinv_node + r0 + 16 * threadIdx.x);
    if (threadIdx.x < (nr + 31) / 32) prefetch_l2(S.vpos + r0 + 32 * threadIdx.x);
  }
  slab_product<DIM, kSlabBatch>(S, s, z, sm, acc);
  for (uint32_t i = threadIdx.x; i < DIM * nr; i += kSlabThreads) {
    const uint32_t r = r0 + i / DIM;
    const double sc = slab_row_sum<DIM>(sm, S.vpos[r], (int)(i % DIM));
    const int64_t g = (int64_t)DIM * r0 + i;
    const double zg = __ldg(z + g);
    const double zo = zold != nullptr ? zold[g] : 0.0;  // nullptr: z_0 = 0 (second sweep of a solve)
    znew[g] = zg + c1 * (zg - zo) + c2 * (bd[g] - dinv_node[r] * sc);
  }
}

// The same sweep with the matrix slices staged by bulk copies (slab_product_tma)
template <int DIM>
__global__ void __launch_bounds__(kSlabThreads)
    fs_slab_sweep_tma_kernel(SlabView S, uint32_t win_doubles, int64_t max_entries, const double *__restrict__ dinv_node,
                             const double *__restrict__ bd, const double *__restrict__ z, const double *__restrict__ zold,
                             double *__restrict__ znew, double c1, double c2) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int s = blockIdx.x;
  double acc[DIM];
  slab_product_tma<DIM, kSlabBatch>(S, s, z, smraw, win_doubles, max_entries, acc);
  const double *sm = reinterpret_cast<const double *>(smraw + kSlabTmaHeader);
  const uint32_t r0 = S.slab_row[s], nr = S.slab_row[s + 1] - r0;
  for (uint32_t i = threadIdx.x; i < DIM * nr; i += kSlabThreads) {
    const uint32_t r = r0 + i / DIM;
    const double sc = slab_row_sum<DIM>(sm, S.vpos[r], (int)(i % DIM));
    const int64_t g = (int64_t)DIM * r0 + i;
    const double zg = __ldg(z + g);
    const double zo = zold != nullptr ? zold[g] : 0.0;
    znew[g] = zg + c1 * (zg - zo) + c2 * (bd[g] - dinv_node[r] * sc);
  }
}

// first sweep on F with zero initial guess: bd = Dinv .* b, z1 = bd / theta
template <int DIM>
__global__ void fs_cheb_first_kernel(int64_t n_u, const double *__restrict__ dinv_node, const double *__restrict__ b,
                                     double inv_theta, double *__restrict__ bd, double *__restrict__ z1) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_u) {
    const double v = dinv_node[i / DIM] * b[i];
    bd[i] = v;
    z1[i] = v * inv_theta;
  }
}

// ---------------------------------------------------------------------------
// A01 (velocity rows x pressure columns, reference system_matrix.block(0,1)) in
// the same slabs, in NODE-BLOCK form: the dim rows of a velocity node have the
// same pattern, so an entry is (pressure column, dim values) -- 8*dim + 2 B
// instead of dim * 12 B in CSR.  The nodes of a slab are sorted by the length
// of their rows (a vertex node couples to ~15 pressure vertices, an edge node
// to ~7) and stored as ELL slices of 32 nodes: thread t of the CTA owns the
// t-th node in that order and streams, per entry, dim coalesced value loads and
// a 16-bit index into the slab's PRESSURE window, which is staged in shared
// memory like the velocity window.  `perm` maps the sorted position back to the
// natural node of the slab so that results are handed over through shared
// memory and written coalesced.
// ---------------------------------------------------------------------------
#ifndef NSB_G_BATCH
#define NSB_G_BATCH 2
#endif
constexpr int kGSlabBatch = NSB_G_BATCH;

struct GSlabView {
  const uint32_t *pwin_ptr;   // n_slabs+1
  const uint32_t *pwin_list;  // pressure columns, ascending inside a slab
  const int64_t *slice_ptr;   // n_slabs*kSlabSlices+1, in entries (multiples of 32); values at dim * offset
  const double *val;          // per slice and entry step k: dim x 32 values, component-major
  const uint16_t *idx;
  const uint16_t *perm;       // per node (slab-major, sorted position) -> natural local node of the slab
};

struct GSlabHost {
  std::vector<uint32_t> pwin_ptr, pwin_list, src;
  std::vector<int64_t> slice_ptr;
  std::vector<uint16_t> idx, perm;
  uint32_t max_window = 0;
  int64_t nnz = 0;
};

struct GSlabDev {
  int64_t nnz = 0, padded = 0;  // scalar values stored / incl. padding
  uint32_t max_window = 0;
  DevBuf<uint32_t> pwin_ptr, pwin_list, src;
  DevBuf<int64_t> slice_ptr;
  DevBuf<double> val;
  DevBuf<uint16_t> idx, perm;
  bool have = false;
  GSlabView view() const { return {pwin_ptr.p, pwin_list.p, slice_ptr.p, val.p, idx.p, perm.p}; }
};

// position of value (step k, component c, lane) of a slice that starts at entry offset `base`
__host__ __device__ inline int64_t gslab_val_pos(int dim, int64_t base, int k, int c, int lane) {
  return (int64_t)dim * base + ((int64_t)k * dim + c) * 32 + lane;
}

// rp/ci: CSR pattern of A01 (one row per velocity dof, dim rows per node, in node order)
inline GSlabHost build_gslabs(int dim, const std::vector<uint32_t> &slab_row, const int64_t *rp, const uint32_t *ci) {
  GSlabHost H;
  const int64_t ns = (int64_t)slab_row.size() - 1;
  const int64_t n_nodes = slab_row.back();
  H.nnz = rp[(int64_t)dim * n_nodes];
  if (H.nnz >= (int64_t)kSlabPad) throw ArgError("slab storage: more than 2^32-2 stored entries per rank");
  for (int64_t a = 0; a < n_nodes; ++a)
    for (int c = 1; c < dim; ++c)
      if (rp[dim * a + c + 1] - rp[dim * a + c] != rp[dim * a + 1] - rp[dim * a])
        throw StructError("A01: the rows of a velocity node do not have the same pattern");
  std::vector<std::vector<uint32_t>> wins((size_t)ns);
  std::vector<int64_t> slice_len((size_t)ns * kSlabSlices, 0);
  H.perm.assign((size_t)n_nodes, 0);
  uint32_t maxw = 0;
  int gerr = 0;  // (exceptions must not leave an OpenMP region)
#pragma omp parallel for schedule(dynamic, 64) reduction(max : maxw, gerr)
  for (int64_t s = 0; s < ns; ++s) {
    const int64_t a0 = slab_row[s], na = (int64_t)slab_row[s + 1] - a0;
    if (na > kSlabThreads) {
      gerr = 1;
      continue;
    }
    std::vector<uint32_t> &w = wins[s];
    for (int64_t a = a0; a < a0 + na; ++a) w.insert(w.end(), ci + rp[dim * a], ci + rp[dim * a + 1]);
    std::sort(w.begin(), w.end());
    w.erase(std::unique(w.begin(), w.end()), w.end());
    if (w.size() > 65535) {
      gerr = 2;
      continue;
    }
    maxw = std::max(maxw, (uint32_t)w.size());
    std::vector<uint16_t> order((size_t)na);
    for (int64_t i = 0; i < na; ++i) order[i] = (uint16_t)i;
    auto len = [&](int64_t a) { return rp[dim * a + 1] - rp[dim * a]; };
    std::stable_sort(order.begin(), order.end(), [&](uint16_t x, uint16_t y) { return len(a0 + x) > len(a0 + y); });
    for (int64_t i = 0; i < na; ++i) H.perm[(size_t)(a0 + i)] = order[i];
    for (int64_t wv = 0; wv * 32 < na; ++wv) slice_len[s * kSlabSlices + wv] = 32 * len(a0 + order[wv * 32]);
  }
  if (gerr == 1) throw StructError("slab storage: more nodes than threads in a slab");
  if (gerr == 2) throw StructError("slab storage: pressure window exceeds 16-bit indices");
  H.max_window = maxw;
  H.pwin_ptr.assign((size_t)ns + 1, 0);
  for (int64_t s = 0; s < ns; ++s) H.pwin_ptr[s + 1] = H.pwin_ptr[s] + (uint32_t)wins[s].size();
  H.pwin_list.resize(H.pwin_ptr.back());
  H.slice_ptr.assign((size_t)ns * kSlabSlices + 1, 0);
  for (size_t i = 0; i < slice_len.size(); ++i) H.slice_ptr[i + 1] = H.slice_ptr[i] + slice_len[i];
  H.idx.assign((size_t)H.slice_ptr.back(), 0);
  H.src.assign((size_t)H.slice_ptr.back() * dim, kSlabPad);
#pragma omp parallel for schedule(dynamic, 64) reduction(max : gerr)
  for (int64_t s = 0; s < ns; ++s) {
    const std::vector<uint32_t> &w = wins[s];
    std::copy(w.begin(), w.end(), H.pwin_list.begin() + H.pwin_ptr[s]);
    const int64_t a0 = slab_row[s], na = (int64_t)slab_row[s + 1] - a0;
    for (int64_t i = 0; i < na; ++i) {
      const int64_t a = a0 + H.perm[(size_t)(a0 + i)];
      const int64_t base = H.slice_ptr[s * kSlabSlices + i / 32];
      const int lane = (int)(i % 32);
      const int64_t r0 = rp[dim * a], n = rp[dim * a + 1] - r0;
      for (int64_t k = 0; k < n; ++k) {
        H.idx[(size_t)(base + 32 * k + lane)] = (uint16_t)(std::lower_bound(w.begin(), w.end(), ci[r0 + k]) - w.begin());
        for (int c = 0; c < dim; ++c) {
          if (ci[rp[dim * a + c] + k] != ci[r0 + k]) gerr = 3;
          H.src[(size_t)gslab_val_pos(dim, base, (int)k, c, lane)] = (uint32_t)(rp[dim * a + c] + k);
        }
      }
    }
  }
  if (gerr == 3) throw StructError("A01: the rows of a velocity node do not have the same pattern");
  return H;
}

inline void upload_gslabs(const GSlabHost &H, GSlabDev &D, cudaStream_t s, int64_t *bytes) {
  D.nnz = H.nnz;
  D.padded = (int64_t)H.src.size();
  D.max_window = H.max_window;
  D.pwin_ptr.upload(H.pwin_ptr.data(), H.pwin_ptr.size(), s, bytes);
  D.pwin_list.upload(H.pwin_list.data(), H.pwin_list.size(), s, bytes);
  D.src.upload(H.src.data(), H.src.size(), s, bytes);
  D.slice_ptr.upload(H.slice_ptr.data(), H.slice_ptr.size(), s, bytes);
  D.idx.upload(H.idx.data(), H.idx.size(), s, bytes);
  D.perm.upload(H.perm.data(), H.perm.size(), s, bytes);
  D.val.alloc((size_t)D.padded, bytes);
  D.val.zero(s);
  NSB_CUDA(cudaStreamSynchronize(s));
  D.have = true;
}

// smo[dim * natural local node + c] = (A01 xp)[dof] for the rows of slab s.  smp: pressure window,
// smo: dim*nr doubles.  Ends with a __syncthreads().
template <int DIM>
__device__ __forceinline__ void slab_g_product(const GSlabView &G, int s, uint32_t r0, uint32_t nr,
                                               const double *__restrict__ xp, double *smp, double *smo) {
  const int t = threadIdx.x;
  const uint32_t w0 = G.pwin_ptr[s], nw = G.pwin_ptr[s + 1] - w0;
  for (uint32_t i = t; i < nw; i += kSlabThreads) smp[i] = __ldg(xp + __ldg(G.pwin_list + w0 + i));
  const int64_t sl = (int64_t)s * kSlabSlices + (t >> 5);
  const int64_t base = G.slice_ptr[sl];
  const int W = (int)((G.slice_ptr[sl + 1] - base) >> 5);  // warp-uniform; 0 for warps beyond the last node
  const double *__restrict__ v = G.val + (int64_t)DIM * base + (t & 31);
  const uint16_t *__restrict__ ix = G.idx + base + (t & 31);
  double acc[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) acc[c] = 0.0;
  __syncthreads();
  constexpr int B = kGSlabBatch;
  if (NSB_G_PF > 0) {  // first batches: DIM value lines (256 B per component and step) + 64 B of indices per step
    for (int q = (t & 31); q < 2 * DIM * NSB_G_PF * B; q += 32)
      if (q / (2 * DIM) < W) prefetch_l2(v - (t & 31) + 16 * q);
  }
  for (int k = 0; k < W; k += B) {
    double a[B][DIM];
    unsigned j[B];
    if (NSB_G_PF > 0) {
      const int k1 = k + NSB_G_PF * B;
      for (int q = (t & 31); q < 2 * DIM * B; q += 32)
        if (k1 + q / (2 * DIM) < W) prefetch_l2(v - (t & 31) + 32 * DIM * k1 + 16 * q);
      if ((t & 31) == 31 && k1 < W) prefetch_l2(ix - (t & 31) + 32 * k1);
    }
#pragma unroll
    for (int u = 0; u < B; ++u) {
      j[u] = k + u < W ? __ldcs(ix + 32 * (k + u)) : 0u;
#pragma unroll
      for (int c = 0; c < DIM; ++c) a[u][c] = k + u < W ? __ldcs(v + 32 * ((k + u) * DIM + c)) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < B; ++u) {
      const double p = smp[j[u]];
#pragma unroll
      for (int c = 0; c < DIM; ++c) acc[c] += a[u][c] * p;
    }
  }
  if ((uint32_t)t < nr) {
    const uint32_t a = __ldg(G.perm + r0 + t);
#pragma unroll
    for (int c = 0; c < DIM; ++c) smo[DIM * a + c] = acc[c];
  }
  __syncthreads();
}

// velocity rows of the block product:
//   MODE 0: y_u = F x_u + A01 x_p     MODE 3: y = d .* (F x)   (power iteration on D^-1 F; G unused)
// shared memory: [velocity window / partial sums][dim*256 A01 results][pressure window]
template <int DIM, int MODE>
__global__ void __launch_bounds__(kSlabThreads) fs_slab_apply_kernel(SlabView S, GSlabView G, uint32_t win_doubles,
                                                                     const double *__restrict__ xu,
                                                                     const double *__restrict__ xp,
                                                                     const double *__restrict__ d,
                                                                     double *__restrict__ y) {
  extern __shared__ double sm[];
  const int s = blockIdx.x;
  const uint32_t r0 = S.slab_row[s], nr = S.slab_row[s + 1] - r0;
  double *smo = sm + win_doubles, *smp = smo + DIM * kSlabThreads;
  if (MODE == 0) slab_g_product<DIM>(G, s, r0, nr, xp, smp, smo);
  double acc[DIM];
  slab_product<DIM, kSlabBatch>(S, s, xu, sm, acc);
  for (uint32_t i = threadIdx.x; i < DIM * nr; i += kSlabThreads) {
    double sc = slab_row_sum<DIM>(sm, S.vpos[r0 + i / DIM], (int)(i % DIM));
    const int64_t g = (int64_t)DIM * r0 + i;
    if (MODE == 0) sc += smo[i];
    y[g] = MODE == 3 ? d[g] * sc : sc;
  }
}

// y = w - d .* (A01 xp)   (dst0 = vec0 - Di .* (Bt dst1), reference :992-994) over the velocity rows;
// w == nullptr: y = A01 xp   (Bt dst1 of PreconditionAYosida::vmult, reference :1047)
template <int DIM>
__global__ void __launch_bounds__(kSlabThreads) g_slab_apply_kernel(SlabView S, GSlabView G,
                                                                    const double *__restrict__ xp,
                                                                    const double *__restrict__ w,
                                                                    const double *__restrict__ d,
                                                                    double *__restrict__ y) {
  extern __shared__ double sm[];
  const int s = blockIdx.x;
  const uint32_t r0 = S.slab_row[s], nr = S.slab_row[s + 1] - r0;
  double *smo = sm, *smp = sm + DIM * kSlabThreads;
  if (NSB_G_PF > 0 && w != nullptr)
    for (uint32_t i = 16 * threadIdx.x; i < DIM * nr; i += 16 * kSlabThreads) {
      prefetch_l2(w + (int64_t)DIM * r0 + i);
      prefetch_l2(d + (int64_t)DIM * r0 + i);
    }
  slab_g_product<DIM>(G, s, r0, nr, xp, smp, smo);
  for (uint32_t i = threadIdx.x; i < DIM * nr; i += kSlabThreads) {
    const int64_t g = (int64_t)DIM * r0 + i;
    y[g] = w != nullptr ? w[g] - d[g] * smo[i] : smo[i];
  }
}

}  // namespace nsb
