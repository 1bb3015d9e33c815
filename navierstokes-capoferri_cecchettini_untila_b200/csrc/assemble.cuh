// Cell-loop assembly of the semi-implicit Oseen system for Taylor-Hood P2/P1
// (replaces reference src/NavierStokes.cpp:154-156, 164-285 and the Dirichlet
// application :326-328).
//
// One warp per cell.  The quadrature tables live in shared memory; the cell is
// affine, so J^{-T} and |det J| are per-cell constants held in registers.  Only
// the structurally non-zero parts of the 34x34 (3D) / 15x15 (2D) element matrix
// are formed:
//   scalar block  A_s(a,b) = |J| [ M^(a,b)/dt + sum_q w_q ( nu G_a.G_b + phi_a (u_q.G_b) ) ]
//                 replicated on the dim diagonal component positions of A00,
//   Bt(a,c;k) = -|J| sum_d Jinv[d][c] D^(a,k,d)   -> A01 and (transposed) A10,
//   rhs(a,c)  = |J|/dt sum_n M^(a,n) U_n[c].
// A00 = F_s (x) I_dim exactly (also after the Dirichlet rows, which hit all
// components of a node alike), so only the scalar node-level matrix F_s is
// stored and assembled: dim^2 times fewer bytes than the canonical block the
// reference keeps; the canonical values are materialised on request
// (nsb_get_matrix_values).  Values are scatter-added with red.global.add.f64
// into precomputed CSR positions: per cell 16-bit "slots" = rank of the column
// node inside the row node's adjacency list.
#pragma once
#include "common.cuh"

namespace nsb {

struct AsmArgs {
  int64_t n_cells;
  const double *xyz;            // n_verts*dim
  const uint32_t *cell_verts;   // n_cells*(dim+1)
  const uint32_t *cell_nodes;   // n_cells*NN
  const uint32_t *cell_pverts;  // n_cells*NV
  const uint16_t *slot00;       // n_cells*NN*NN
  const uint16_t *slot01;       // n_cells*NN*NV
  const uint16_t *slot10;       // n_cells*NV*NN
  const int64_t *nptr, *rowptr01, *rowptr10;  // nptr: node-level row offsets of F_s
  double *fs_val, *val01, *val10;
  // ownership (multi-GPU): rows of local nodes >= n_own_nodes and of pressure
  // vertices outside [p_begin, p_begin + n_p_own) belong to another rank
  uint32_t n_own_nodes, p_begin, n_p_own;
  double *rhs;
  const double *sol;  // previous-step solution (velocity part is read)
  const FeTables *fe;
  double inv_dt, nu;
};

// ---------------------------------------------------------------------------
// setup: scatter slots and diagonal positions by binary search in the
// canonical pattern; `err` is raised when an expected column is absent or the
// Taylor-Hood structure (row of dim*A has dim entries per neighbour node) fails.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int64_t find_col(const uint32_t *colind, int64_t b, int64_t e, uint32_t col) {
  int64_t lo = b, hi = e;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (colind[mid] < col)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (lo < e && colind[lo] == col) ? lo : -1;
}

template <int DIM>
__global__ void build_slots_kernel(int64_t n_cells, const uint32_t *__restrict__ cell_nodes,
                                   const uint32_t *__restrict__ cell_pverts, CsrView fs, CsrView a01, CsrView a10,
                                   uint16_t *slot00, uint16_t *slot01, uint16_t *slot10, uint32_t p_begin,
                                   int *err) {
  constexpr int NV = DIM + 1, NN = DIM == 2 ? 6 : 10, PER = NN * NN + 2 * NN * NV;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cells * PER) return;
  const int64_t cell = t / PER;
  int e = (int)(t % PER);
  const uint32_t *nodes = cell_nodes + cell * NN, *pv = cell_pverts + cell * NV;
  if (e < NN * NN) {
    const int a = e / NN, b = e % NN;
    const int64_t row = nodes[a];
    if (row >= fs.n_rows) return;  // row owned by another rank
    const int64_t rb = fs.rowptr[row], re = fs.rowptr[row + 1];
    const int64_t pos = find_col(fs.colind, rb, re, nodes[b]);
    if (pos < 0 || pos - rb > 65535) atomicExch(err, 1);
    slot00[cell * NN * NN + e] = (uint16_t)(pos - rb);
    return;
  }
  e -= NN * NN;
  if (e < NN * NV) {
    const int a = e / NV, k = e % NV;
    const int64_t row = (int64_t)DIM * nodes[a];
    if (row >= a01.n_rows) return;
    const int64_t rb = a01.rowptr[row], re = a01.rowptr[row + 1];
    const int64_t pos = find_col(a01.colind, rb, re, pv[k]);
    if (pos < 0 || pos - rb > 65535) atomicExch(err, 2);
    slot01[cell * NN * NV + e] = (uint16_t)(pos - rb);
    return;
  }
  e -= NN * NV;
  {
    const int k = e / NN, a = e % NN;
    const int64_t row = (int64_t)pv[k] - p_begin;  // local row of an owned pressure vertex
    if (row < 0 || row >= a10.n_rows) return;
    const int64_t rb = a10.rowptr[row], re = a10.rowptr[row + 1];
    const int64_t pos = find_col(a10.colind, rb, re, DIM * nodes[a]);
    const int64_t s = (pos - rb) / DIM;
    if (pos < 0 || (pos - rb) % DIM != 0 || s > 65535) atomicExch(err, 3);
    slot10[cell * NV * NN + e] = (uint16_t)s;
  }
}

__global__ void diag_positions_kernel(CsrView A, int64_t *diagpos, int *err) {  // A: F_s (node level) or S
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n_rows) return;
  const int64_t pos = find_col(A.colind, A.rowptr[i], A.rowptr[i + 1], (uint32_t)i);
  if (pos < 0) atomicExch(err, 4);
  diagpos[i] = pos;
}

// cell_dofs (deal.II FESystem order) -> P2 node ids and pressure vertex ids
template <int DIM>
__global__ void split_cell_dofs_kernel(int64_t n_cells, const uint32_t *__restrict__ cell_dofs, uint32_t n_u,
                                       uint32_t n_p, uint32_t *cell_nodes, uint32_t *cell_pverts, int *err) {
  constexpr int NV = DIM + 1, NL = DIM == 2 ? 3 : 6, NN = NV + NL, DPC = DIM * NN + NV;
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  const uint32_t *d = cell_dofs + cell * DPC;
  bool bad = false;
  for (int a = 0; a < NV; ++a) {
    const uint32_t u0 = d[a * (DIM + 1)];
    for (int c = 0; c < DIM; ++c) bad |= d[a * (DIM + 1) + c] != u0 + c;
    bad |= (u0 % DIM) != 0 || u0 + DIM > n_u;
    const uint32_t p = d[a * (DIM + 1) + DIM];
    bad |= p < n_u || p - n_u >= n_p;
    cell_nodes[cell * NN + a] = u0 / DIM;
    cell_pverts[cell * NV + a] = p - n_u;
  }
  for (int l = 0; l < NL; ++l) {
    const uint32_t u0 = d[NV * (DIM + 1) + l * DIM];
    for (int c = 0; c < DIM; ++c) bad |= d[NV * (DIM + 1) + l * DIM + c] != u0 + c;
    bad |= (u0 % DIM) != 0 || u0 + DIM > n_u;
    cell_nodes[cell * NN + NV + l] = u0 / DIM;
  }
  if (bad) atomicExch(err, 5);
}

// node-level adjacency -> canonical A00 pattern (nodes (x) ones(dim,dim))
template <int DIM>
__global__ void expand_node_pattern_kernel(int64_t n_nodes, const int64_t *__restrict__ nptr,
                                           const uint32_t *__restrict__ ncol, int64_t *rowptr, uint32_t *colind) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_nodes) return;
  const int64_t b = nptr[w], len = nptr[w + 1] - b;
  if (lane < DIM) rowptr[DIM * w + lane] = DIM * DIM * b + lane * DIM * len;
  if (w == n_nodes - 1 && lane == 0) rowptr[DIM * n_nodes] = DIM * DIM * nptr[n_nodes];
  for (int c = 0; c < DIM; ++c) {
    uint32_t *o = colind + DIM * DIM * b + c * DIM * len;
    for (int64_t k = lane; k < DIM * len; k += 32) o[k] = DIM * ncol[b + k / DIM] + (uint32_t)(k % DIM);
  }
}

// ---------------------------------------------------------------------------
// the cell loop
// ---------------------------------------------------------------------------
constexpr int kAsmWarps = 8;

// The scalar velocity block needs no quadrature loop at run time.  With G_a = J^{-T} grad_hat(phi_a) and
// u_q = sum_n phi_n(q) U_n (reference :175, :197-208):
//   sum_q w_q G_a.G_b               = sum_{d,e} Q_de khat[de](a,b),        Q = J^{-1} J^{-T}
//   sum_q w_q phi_a (u_q . G_b)     = sum_{n,d} chat[nd](a,b) Uh_nd,       Uh_nd = sum_c J^{-1}[d][c] U_n[c]
// khat / chat are contractions of the SAME quadrature table the reference uses (fe_tables.h), so its
// under-integration is reproduced; 9 + 30 conflict-free shared-memory reads per (a,b) pair replace the
// 14 x 12 of the quadrature loop, which made the kernel shared-memory-bandwidth bound (13 ms at 2.25 M cells).
template <int DIM>
__global__ void __launch_bounds__(kAsmWarps * 32) assemble_cells_kernel(AsmArgs A) {
  constexpr int NV = DIM + 1, NN = DIM == 2 ? 6 : 10, NP = NN * NN, ND = NN * DIM;
  __shared__ double s_mhat[NN][NN], s_dhat[NN][NV][DIM], s_khat[DIM * DIM][NP], s_chat[ND][NP];
  __shared__ double s_U[kAsmWarps][NN][DIM];
  __shared__ double s_Uh[kAsmWarps][ND];
  __shared__ int64_t s_rp00[kAsmWarps][NN], s_rp01[kAsmWarps][NN], s_rp10[kAsmWarps][NV];
  __shared__ int s_len01[kAsmWarps][NN];
  __shared__ uint32_t s_node[kAsmWarps][NN];

  for (int i = threadIdx.x; i < NN * NN; i += blockDim.x) s_mhat[i / NN][i % NN] = A.fe->mhat[i / NN][i % NN];
  for (int i = threadIdx.x; i < NN * NV * DIM; i += blockDim.x)
    s_dhat[i / (NV * DIM)][(i / DIM) % NV][i % DIM] = A.fe->dhat[i / (NV * DIM)][(i / DIM) % NV][i % DIM];
  for (int i = threadIdx.x; i < DIM * DIM * NP; i += blockDim.x) s_khat[i / NP][i % NP] = A.fe->khat[i / NP][i % NP];
  for (int i = threadIdx.x; i < ND * NP; i += blockDim.x) s_chat[i / NP][i % NP] = A.fe->chat[i / NP][i % NP];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t cell = (int64_t)blockIdx.x * kAsmWarps + warp; cell < A.n_cells;
       cell += (int64_t)gridDim.x * kAsmWarps) {
    // ---- affine geometry (all lanes read the same addresses: broadcast) ----
    const uint32_t *cv = A.cell_verts + cell * NV;
    double X[NV][DIM];
#pragma unroll
    for (int a = 0; a < NV; ++a) {
      const uint32_t v = __ldg(cv + a);
#pragma unroll
      for (int r = 0; r < DIM; ++r) X[a][r] = __ldg(A.xyz + (size_t)v * DIM + r);
    }
    double Ji[DIM][DIM], det;  // Ji = J^{-1}, J[r][a] = X[a+1][r]-X[0][r]
    if constexpr (DIM == 2) {
      const double j00 = X[1][0] - X[0][0], j01 = X[2][0] - X[0][0];
      const double j10 = X[1][1] - X[0][1], j11 = X[2][1] - X[0][1];
      det = j00 * j11 - j01 * j10;
      const double id = 1.0 / det;
      Ji[0][0] = j11 * id;
      Ji[0][1] = -j01 * id;
      Ji[1][0] = -j10 * id;
      Ji[1][1] = j00 * id;
    } else {
      double J[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int r = 0; r < 3; ++r) J[r][a] = X[a + 1][r] - X[0][r];
      const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
      const double id = 1.0 / det;
      Ji[0][0] = c00 * id;
      Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
      Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
      Ji[1][0] = c01 * id;
      Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
      Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
      Ji[2][0] = c02 * id;
      Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
      Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    }
    const double adet = fabs(det);

    // ---- dof ids and row bases ----
    if (lane < NN) {
      const uint32_t node = __ldg(A.cell_nodes + cell * NN + lane);
      s_node[warp][lane] = node;
      if (node < A.n_own_nodes) {  // rows of nodes this rank owns
        s_rp00[warp][lane] = __ldg(A.nptr + node);
        const int64_t q0 = __ldg(A.rowptr01 + (int64_t)DIM * node), q1 = __ldg(A.rowptr01 + (int64_t)DIM * node + 1);
        s_rp01[warp][lane] = q0;
        s_len01[warp][lane] = (int)(q1 - q0);
      } else
        s_rp00[warp][lane] = -1;
    } else if (lane >= 16 && lane < 16 + NV) {
      const uint32_t pv = __ldg(A.cell_pverts + cell * NV + (lane - 16)) - A.p_begin;  // wraps when below p_begin
      s_rp10[warp][lane - 16] = pv < A.n_p_own ? __ldg(A.rowptr10 + pv) : -1;
    }
    __syncwarp();
    // previous-step velocity at the cell's nodes (reference :175), as is and pulled back to the reference cell
    for (int e = lane; e < ND; e += 32) {
      const int n = e / DIM, c = e % DIM;
      s_U[warp][n][c] = __ldg(A.sol + (size_t)DIM * s_node[warp][n] + c);
    }
    __syncwarp();
    for (int e = lane; e < ND; e += 32) {
      const int n = e / DIM, d = e % DIM;
      double u = 0;
#pragma unroll
      for (int c = 0; c < DIM; ++c) u += Ji[d][c] * s_U[warp][n][c];
      s_Uh[warp][e] = u;
    }
    double Q[DIM][DIM];  // J^{-1} J^{-T}
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int e = 0; e < DIM; ++e) {
        double q = 0;
#pragma unroll
        for (int c = 0; c < DIM; ++c) q += Ji[d][c] * Ji[e][c];
        Q[d][e] = q;
      }
    __syncwarp();

    // ---- scalar velocity block -> F_s ----
    const uint16_t *sl00 = A.slot00 + cell * NP;
    for (int p = lane; p < NP; p += 32) {
      const int a = p / NN;
      double stiff = 0, conv = 0;
#pragma unroll
      for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int e = 0; e < DIM; ++e) stiff += Q[d][e] * s_khat[d * DIM + e][p];
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) conv += s_chat[nd][p] * s_Uh[warp][nd];
      const double v = adet * (A.nu * stiff + conv + s_mhat[a][p % NN] * A.inv_dt);
      if (s_rp00[warp][a] >= 0) atomicAdd(A.fs_val + s_rp00[warp][a] + sl00[p], v);
    }
    // ---- pressure-velocity coupling -> A01 and A10 (reference :222-229) ----
    const uint16_t *sl01 = A.slot01 + cell * (NN * NV), *sl10 = A.slot10 + cell * (NV * NN);
    for (int e = lane; e < NN * DIM * NV; e += 32) {
      const int a = e / (DIM * NV), c = (e / NV) % DIM, k = e % NV;
      double s = 0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) s += Ji[d][c] * s_dhat[a][k][d];
      const double v = -adet * s;
      if (s_rp00[warp][a] >= 0) {
        const int64_t pos = s_rp01[warp][a] + (int64_t)c * s_len01[warp][a] + sl01[a * NV + k];
        atomicAdd(A.val01 + pos, v);
      }
      if (s_rp10[warp][k] >= 0) atomicAdd(A.val10 + s_rp10[warp][k] + (int64_t)DIM * sl10[k * NN + a] + c, v);
    }
    // ---- right-hand side: (u^n, v)/dt, forcing term f == 0 (reference :241-248) ----
    for (int e = lane; e < NN * DIM; e += 32) {
      const int a = e / DIM, c = e % DIM;
      double s = 0;
#pragma unroll
      for (int n = 0; n < NN; ++n) s += s_mhat[a][n] * s_U[warp][n][c];
      if (s_rp00[warp][a] >= 0) atomicAdd(A.rhs + (size_t)DIM * s_node[warp][a] + c, adet * A.inv_dt * s);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// Dirichlet rows (MatrixTools::apply_boundary_values, eliminate_columns=false,
// reference :326-328).  One warp per constrained dof.
// ---------------------------------------------------------------------------
__global__ void first_diag_kernel(const double *__restrict__ fs_val, const int64_t *__restrict__ diagpos,
                                  int64_t n_nodes, double *first_diag) {
  // "first non-zero diagonal entry in the local range": a serial scan that in
  // practice stops at row 0.
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double d = 1.0;
    for (int64_t i = 0; i < n_nodes; ++i) {
      const double v = fs_val[diagpos[i]];
      if (v != 0.0) {
        d = fabs(v);
        break;
      }
    }
    *first_diag = d;
  }
}

// bc_nodes: constrained nodes (all dim components of a node are constrained
// together in the reference, :300-324); vals: dim values per node.
template <int DIM>
__global__ void apply_dirichlet_kernel(int64_t n_bc_nodes, const uint32_t *__restrict__ bc_nodes,
                                       const double *__restrict__ vals, double factor, CsrView fs, CsrView a01,
                                       const int64_t *__restrict__ diagpos, const double *__restrict__ first_diag,
                                       int mode, double *rhs, double *sol) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_bc_nodes) return;
  const uint32_t A = bc_nodes[w];
  const int64_t dp = diagpos[A];
  double dg = fs.val[dp];
  __syncwarp();
  if (mode == NSB_BCDIAG_FIRST || dg == 0.0) dg = *first_diag;
  for (int64_t k = fs.rowptr[A] + lane; k < fs.rowptr[A + 1]; k += 32) fs.val[k] = (k == dp) ? dg : 0.0;
  for (int64_t k = a01.rowptr[(int64_t)DIM * A] + lane; k < a01.rowptr[(int64_t)DIM * A + DIM]; k += 32)
    a01.val[k] = 0.0;
  if (lane < DIM) {
    const double g = vals[w * DIM + lane] * factor;
    rhs[(int64_t)DIM * A + lane] = g * dg;
    sol[(int64_t)DIM * A + lane] = g;
  }
}

// LUMPED == false: diagonal of the velocity mass matrix, M_AA = sum_cells |J| M^(a,a) (used to choose the
// degree of the Chebyshev polynomial for F, see auto_inner in nsb_capi.cu).
// LUMPED == true: the reference's lumped mass sum_cells sum_q sum_b |phi_a phi_b JxW| per node
// (NavierStokes.cpp:232-236, 252, 284; consumed by PreconditionAYosida as deltat / lumped).
// Geometry only; computed once.
template <int DIM, bool LUMPED>
__global__ void mass_diag_kernel(int64_t n_cells, const double *__restrict__ xyz,
                                 const uint32_t *__restrict__ cell_verts, const uint32_t *__restrict__ cell_nodes,
                                 uint32_t n_own_nodes, const FeTables *__restrict__ fe, double *mdiag) {
  constexpr int NV = DIM + 1, NN = DIM == 2 ? 6 : 10;
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  double X[NV][DIM];
  for (int a = 0; a < NV; ++a)
    for (int r = 0; r < DIM; ++r) X[a][r] = xyz[(size_t)cell_verts[cell * NV + a] * DIM + r];
  double det;
  if constexpr (DIM == 2) {
    det = (X[1][0] - X[0][0]) * (X[2][1] - X[0][1]) - (X[2][0] - X[0][0]) * (X[1][1] - X[0][1]);
  } else {
    double J[3][3];
    for (int a = 0; a < 3; ++a)
      for (int r = 0; r < 3; ++r) J[r][a] = X[a + 1][r] - X[0][r];
    det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
          J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  }
  for (int a = 0; a < NN; ++a) {
    const uint32_t node = cell_nodes[cell * NN + a];
    if (node < n_own_nodes) atomicAdd(mdiag + node, fabs(det) * (LUMPED ? fe->labs[a] : fe->mhat[a][a]));
  }
}

// w[A] = sqrt( F_AA * dt / M_AA ): sum of squares = sum of the diagonal growth ratios
__global__ void diag_ratio_kernel(int64_t n_nodes, const double *__restrict__ fs_val,
                                  const int64_t *__restrict__ diagpos, const double *__restrict__ mdiag, double dt,
                                  double *__restrict__ w) {
  const int64_t A = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (A < n_nodes) w[A] = sqrt(fabs(fs_val[diagpos[A]]) * dt / mdiag[A]);
}

// canonical A00 values from F_s: val[rowptr(dim*A+c) + dim*k + c'] = (c == c') F_s[A,k]
template <int DIM>
__global__ void expand_values_kernel(int64_t n_nodes, const int64_t *__restrict__ nptr,
                                     const double *__restrict__ fs_val, double *__restrict__ val00) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_nodes) return;
  const int64_t b = nptr[w], len = nptr[w + 1] - b;
  for (int c = 0; c < DIM; ++c) {
    double *o = val00 + DIM * DIM * b + c * DIM * len;
    for (int64_t k = lane; k < DIM * len; k += 32) o[k] = (k % DIM == c) ? fs_val[b + k / DIM] : 0.0;
  }
}

// node-level adjacency from a canonical A00 pattern (and structure check)
template <int DIM>
__global__ void compress_pattern_kernel(int64_t n_nodes, const int64_t *__restrict__ rowptr,
                                        const uint32_t *__restrict__ colind, int64_t *nptr, uint32_t *ncol, int *err) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_nodes) return;
  const int64_t rb = rowptr[DIM * w], len = (rowptr[DIM * w + 1] - rb) / DIM;
  bool bad = rb % (DIM * DIM) != 0 || (rowptr[DIM * w + 1] - rb) % DIM != 0;
  for (int c = 1; c < DIM; ++c) bad |= rowptr[DIM * w + c + 1] - rowptr[DIM * w + c] != DIM * len;
  if (lane == 0) nptr[w] = rb / (DIM * DIM);
  if (w == n_nodes - 1 && lane == 0) nptr[n_nodes] = rowptr[DIM * n_nodes] / (DIM * DIM);
  for (int64_t k = lane; k < len; k += 32) {
    const uint32_t c0 = colind[rb + DIM * k];
    bad |= c0 % DIM != 0;
    for (int c = 1; c < DIM; ++c) bad |= colind[rb + DIM * k + c] != c0 + c;
    if (!bad) ncol[rb / (DIM * DIM) + k] = c0 / DIM;
  }
  if (bad) atomicExch(err, 6);
}

}  // namespace nsb
