// Internal types of the CPU oracle, shared by ns_oracle.cpp (the checker) and
// ns_baseline.cpp (the partitioned all-core CPU baseline bench.py times).
// TEST INFRASTRUCTURE ONLY -- see ns_oracle.h (parity unpinned).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <vector>

#include "ns_oracle.h"

typedef std::vector<double> Vec;

struct CsrMat {
  int64_t n_rows = 0, n_cols = 0;
  std::vector<int64_t> rowptr;
  std::vector<uint32_t> colind;
  std::vector<double> val;
  int64_t find(int64_t r, uint32_t c) const {
    auto b = colind.begin() + rowptr[r], e = colind.begin() + rowptr[r + 1];
    auto it = std::lower_bound(b, e, c);
    return (it != e && *it == c) ? it - colind.begin() : -1;
  }
  void vmult(double *y, const double *x) const {  // Epetra row-wise CSR product
    for (int64_t r = 0; r < n_rows; ++r) {
      double s = 0;
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) s += val[k] * x[colind[k]];
      y[r] = s;
    }
  }
};

struct NsoBaseline;  // ns_baseline.cpp

struct Quad {
  std::vector<std::array<double, 3>> pt;
  std::vector<double> w;
};

struct nso {
  int dim, nv, nl, NN, dpc, nq, nqf, rule;
  int64_t n_verts, n_cells;
  std::vector<double> xyz;
  std::vector<uint32_t> cells;
  std::vector<uint32_t> bfaces;
  std::vector<int32_t> bids;
  // numbering
  uint32_t n_u = 0, n_p = 0;
  std::vector<uint32_t> cell_dofs;          // n_cells*dpc, deal.II local order
  std::vector<std::array<double, 3>> support;  // support point of every dof
  std::vector<int> local_comp, local_scalar;   // per local dof: component, scalar shape index
  // boundary faces: (cell, local face, id)
  struct BFace {
    uint32_t cell;
    int lf, id;
  };
  std::vector<BFace> bf;
  // system
  CsrMat A00, A01, A10, S;
  Vec rhs, lumped, solution_owned, solution;
  std::vector<uint32_t> bc_dofs;
  std::vector<double> bc_vals;
  // reference tables
  Quad quad;
  std::vector<double> wface;
  std::vector<double> phi;    // nq*NN
  std::vector<double> dphi;   // nq*NN*3 (reference gradients)
  std::vector<double> psi;    // nq*nv
  // parameters (NavierStokes.hpp:254-256, :306)
  double nu = 1e-3, p_out = 0.0, Diameter = 0.4, deltat = 0.01, alpha = 0.5;
  int inlet_kind = NSO_INLET_PARABOLIC, inlet_sin = 0, bc_diag_mode = 0;
  double U_m = 0.3, H = 0.41, inlet_time = 0.0;
  double outer_rtol = 1e-6, inner_rtol = 1e-2;
  int n_tmp = 30, max_it = 10000, threads = 1;
  NsoBaseline *baseline = nullptr;  // partitioned CPU-baseline mode (ns_baseline.cpp); unused by the checker

  double inlet_value(const double *p, int comp, double t) const;
  double mean_vel(double t) const;
};

// Per-cell FEValues data: vector-valued shape values/gradients through the
// extractors (fe_values[velocity].value / gradient / divergence,
// fe_values[pressure].value), kept as full tensors -- the reference's triple
// loop multiplies them out term by term.
struct CellFE {
  // [q][i][d], [q][i][d][e] (component d, derivative e), [q][i]
  std::vector<double> val, grad, div, pval, JxW;
};


extern "C" {
// reference :171-254, the literal (q, i, j) loop of one cell (ns_oracle.cpp)
void nso_cell_contribution(const nso *o, int64_t c, CellFE &fe, double *cell_matrix, double *cell_rhs,
                           double *cell_lumped);
// reference :297-329: interpolate_boundary_values + apply_boundary_values on the assembled system (ns_oracle.cpp)
void nso_apply_boundary(nso *o, double time);
void nso_baseline_free(nso *o);  // ns_baseline.cpp
}
