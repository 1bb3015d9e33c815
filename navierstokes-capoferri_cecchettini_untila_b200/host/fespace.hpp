// Taylor-Hood P2/P1 DoF numbering, block sparsity pattern, Dirichlet DoF
// lists and obstacle-face lists on a simplex mesh: the immutable inputs of the
// hot path that the reference gets from deal.II in NavierStokes::setup
// (reference src/NavierStokes.cpp:35-41, 65-70, 101-117) and from
// VectorTools::interpolate_boundary_values (:297-324).  Semantics: SURVEY.md
// Appendix A.2, A.3, A.5, A.7.
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

#include "mesh.hpp"

namespace nsb {

// Local entity tables of the reference simplex (deal.II ReferenceCell order,
// SURVEY.md A.2).
extern const int kTriLines[3][2];
extern const int kTetLines[6][2];
extern const int kTetFaces[4][3];

inline int n_p2_nodes(int dim) { return dim == 2 ? 6 : 10; }

struct DofMap {
  int dim = 0;
  uint32_t n_nodes = 0;   // P2 nodes (vertices + edges), first-encounter order
  uint32_t n_pverts = 0;  // P1 nodes (vertices), first-encounter order
  uint32_t n_u = 0, n_p = 0;
  std::vector<uint32_t> cell_nodes;   // n_cells*NN : local vertex a<dim+1, then local lines
  std::vector<uint32_t> cell_pverts;  // n_cells*(dim+1)
  std::vector<uint32_t> cell_dofs;    // n_cells*dpc, deal.II FESystem cell order, global dof ids
  std::vector<uint32_t> vert_node;    // mesh vertex -> P2 node
  std::vector<uint32_t> vert_pvert;   // mesh vertex -> pressure index
  std::vector<double> node_xyz;       // support points, n_nodes*dim
  // sorted (min<<32|max) vertex-pair keys and the node of that edge
  std::vector<uint64_t> edge_keys;
  std::vector<uint32_t> edge_nodes;
  int nn() const { return n_p2_nodes(dim); }
  int dofs_per_cell() const { return dim * nn() + dim + 1; }
  uint32_t edge_node(uint32_t a, uint32_t b) const;  // UINT32_MAX if absent
};

DofMap build_dofmap(const Mesh &m);

// CSR with 32-bit column indices (types::global_dof_index is 32-bit) and 64-bit
// row offsets (C5's A00 has ~8.6e8 entries).
struct Csr {
  uint32_t n_rows = 0, n_cols = 0;
  std::vector<int64_t> rowptr;
  std::vector<uint32_t> colind;
  int64_t nnz() const { return rowptr.empty() ? 0 : rowptr.back(); }
};

struct Patterns {
  Csr nodes;  // node x node adjacency (shares a cell), columns ascending
  Csr a00;    // n_u x n_u canonical: nodes (x) ones(dim,dim)
  Csr a01;    // n_u x n_p (block-local columns)
  Csr a10;    // n_p x n_u
  Csr s;      // n_p x n_p pattern of A10*A01 (for S = B diag(Di) Bt, reference :956)
};

// `expand_a00 = false` leaves a00 empty (device expands from `nodes`).
Patterns build_patterns(const Mesh &m, const DofMap &d, bool expand_a00 = true);

struct BoundaryFace {
  uint32_t cell;
  int local_face;
  int id;
};
// All boundary facets of the mesh (exactly one adjacent cell), id from the
// tagged facets, 0 when untagged.
std::vector<BoundaryFace> boundary_faces(const Mesh &m);

// Velocity Dirichlet DoFs, reference :297-324: first faces with id 3, then ids
// 0, 2 (inlet function) and 4 (zero); later writes win; result sorted by dof.
// `profile(x, comp)` is InletVelocity::value at time factor 1.
struct DirichletSet {
  std::vector<uint32_t> dofs;
  std::vector<double> values;  // profile values (multiply by the time factor)
};
DirichletSet dirichlet_dofs(const Mesh &m, const DofMap &d, const std::vector<BoundaryFace> &bf,
                            const std::function<double(const double *, int)> &profile);

// Faces with boundary id 4 (reference :874-875): cell, outward unit normal,
// face measure.
struct ForceFaces {
  std::vector<uint32_t> cell;
  std::vector<double> normal;   // n*dim
  std::vector<double> measure;  // length / area
};
ForceFaces force_faces(const Mesh &m, const std::vector<BoundaryFace> &bf, int id = 4);

}  // namespace nsb
