// See fespace.hpp.
#include "fespace.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <map>
#include <stdexcept>
#include <unordered_map>

namespace nsb {

const int kTriLines[3][2] = {{0, 1}, {1, 2}, {2, 0}};
const int kTetLines[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
const int kTetFaces[4][3] = {{0, 1, 2}, {1, 0, 3}, {0, 2, 3}, {2, 1, 3}};

namespace {
inline uint64_t ekey(uint32_t a, uint32_t b) {
  return a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a;
}
}  // namespace

uint32_t DofMap::edge_node(uint32_t a, uint32_t b) const {
  const uint64_t k = ekey(a, b);
  auto it = std::lower_bound(edge_keys.begin(), edge_keys.end(), k);
  if (it == edge_keys.end() || *it != k) return UINT32_MAX;
  return edge_nodes[it - edge_keys.begin()];
}

// walk the cells in order; per cell first the not-yet-numbered
// vertices (dim+1 dofs each: u_0..u_{dim-1}, p), then the not-yet-numbered
// lines (dim dofs each); then DoFRenumbering::component_wise with blocks
// {u..u -> 0, p -> 1} (reference :68-70) = a stable partition, so that
// velocity dof = dim*node + c and pressure dof = n_u + (vertex rank).
DofMap build_dofmap(const Mesh &m) {
  DofMap d;
  d.dim = m.dim;
  const int dim = m.dim, nv = dim + 1, nl = dim == 2 ? 3 : 6, NN = nv + nl;
  const size_t nc = m.n_cells();
  d.vert_node.assign(m.n_verts(), UINT32_MAX);
  d.vert_pvert.assign(m.n_verts(), UINT32_MAX);
  d.cell_nodes.resize(nc * NN);
  d.cell_pverts.resize(nc * nv);
  std::unordered_map<uint64_t, uint32_t> edges;
  edges.reserve(nc * (dim == 2 ? 2 : 2));
  uint32_t next_node = 0, next_p = 0;
  for (size_t c = 0; c < nc; ++c) {
    const uint32_t *v = &m.cells[c * nv];
    for (int a = 0; a < nv; ++a) {
      if (d.vert_node[v[a]] == UINT32_MAX) {
        d.vert_node[v[a]] = next_node++;
        d.vert_pvert[v[a]] = next_p++;
      }
      d.cell_nodes[c * NN + a] = d.vert_node[v[a]];
      d.cell_pverts[c * nv + a] = d.vert_pvert[v[a]];
    }
    for (int l = 0; l < nl; ++l) {
      const int *lv = dim == 2 ? kTriLines[l] : kTetLines[l];
      auto ins = edges.emplace(ekey(v[lv[0]], v[lv[1]]), next_node);
      if (ins.second) ++next_node;
      d.cell_nodes[c * NN + nv + l] = ins.first->second;
    }
  }
  d.n_nodes = next_node;
  d.n_pverts = next_p;
  d.n_u = dim * next_node;
  d.n_p = next_p;
  // deal.II cell-local order: per vertex (u_0..u_{dim-1}, p), then per line (u_0..u_{dim-1})
  const int dpc = d.dofs_per_cell();
  d.cell_dofs.resize(nc * dpc);
  for (size_t c = 0; c < nc; ++c) {
    uint32_t *o = &d.cell_dofs[c * dpc];
    for (int a = 0; a < nv; ++a) {
      for (int k = 0; k < dim; ++k) *o++ = dim * d.cell_nodes[c * NN + a] + k;
      *o++ = d.n_u + d.cell_pverts[c * nv + a];
    }
    for (int l = 0; l < nl; ++l)
      for (int k = 0; k < dim; ++k) *o++ = dim * d.cell_nodes[c * NN + nv + l] + k;
  }
  // support points and the sorted edge table
  d.node_xyz.assign((size_t)d.n_nodes * dim, 0.0);
  for (size_t v = 0; v < m.n_verts(); ++v)
    if (d.vert_node[v] != UINT32_MAX)
      for (int r = 0; r < dim; ++r) d.node_xyz[(size_t)d.vert_node[v] * dim + r] = m.xyz[v * dim + r];
  std::vector<std::pair<uint64_t, uint32_t>> es(edges.begin(), edges.end());
  std::sort(es.begin(), es.end());
  d.edge_keys.resize(es.size());
  d.edge_nodes.resize(es.size());
  for (size_t i = 0; i < es.size(); ++i) {
    d.edge_keys[i] = es[i].first;
    d.edge_nodes[i] = es[i].second;
    const uint32_t a = (uint32_t)(es[i].first >> 32), b = (uint32_t)es[i].first;
    for (int r = 0; r < dim; ++r)
      d.node_xyz[(size_t)es[i].second * dim + r] = 0.5 * (m.xyz[(size_t)a * dim + r] + m.xyz[(size_t)b * dim + r]);
  }
  return d;
}

namespace {

// rows -> sorted unique union of `cols_of_cell` over the cells incident to the
// row entity.  inc_ptr/inc = entity -> cells incidence.
Csr adjacency(uint32_t n_rows, uint32_t n_cols, const std::vector<int64_t> &inc_ptr,
              const std::vector<uint32_t> &inc, const std::vector<uint32_t> &cell_cols, int per_cell) {
  Csr A;
  A.n_rows = n_rows;
  A.n_cols = n_cols;
  A.rowptr.assign((size_t)n_rows + 1, 0);
  std::vector<uint32_t> len(n_rows);
#pragma omp parallel
  {
    std::vector<uint32_t> buf;
#pragma omp for schedule(static)
    for (int64_t r = 0; r < (int64_t)n_rows; ++r) {
      buf.clear();
      for (int64_t k = inc_ptr[r]; k < inc_ptr[r + 1]; ++k) {
        const uint32_t *cc = &cell_cols[(size_t)inc[k] * per_cell];
        buf.insert(buf.end(), cc, cc + per_cell);
      }
      std::sort(buf.begin(), buf.end());
      len[r] = (uint32_t)(std::unique(buf.begin(), buf.end()) - buf.begin());
    }
  }
  for (uint32_t r = 0; r < n_rows; ++r) A.rowptr[r + 1] = A.rowptr[r] + len[r];
  A.colind.resize((size_t)A.rowptr[n_rows]);
#pragma omp parallel
  {
    std::vector<uint32_t> buf;
#pragma omp for schedule(static)
    for (int64_t r = 0; r < (int64_t)n_rows; ++r) {
      buf.clear();
      for (int64_t k = inc_ptr[r]; k < inc_ptr[r + 1]; ++k) {
        const uint32_t *cc = &cell_cols[(size_t)inc[k] * per_cell];
        buf.insert(buf.end(), cc, cc + per_cell);
      }
      std::sort(buf.begin(), buf.end());
      buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
      std::copy(buf.begin(), buf.end(), A.colind.begin() + A.rowptr[r]);
    }
  }
  return A;
}

void incidence(uint32_t n_ent, const std::vector<uint32_t> &cell_ent, int per_cell, std::vector<int64_t> &ptr,
               std::vector<uint32_t> &inc) {
  ptr.assign((size_t)n_ent + 1, 0);
  for (uint32_t e : cell_ent) ++ptr[e + 1];
  for (uint32_t e = 0; e < n_ent; ++e) ptr[e + 1] += ptr[e];
  inc.resize(cell_ent.size());
  std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
  const size_t nc = cell_ent.size() / per_cell;
  for (size_t c = 0; c < nc; ++c)
    for (int a = 0; a < per_cell; ++a) inc[fill[cell_ent[c * per_cell + a]]++] = (uint32_t)c;
}

}  // namespace

// SURVEY.md A.5 / reference :101-117: every component pair couples except
// (p,p); DoFTools::make_sparsity_pattern inserts, per cell, all (i,j) pairs of
// coupled components.  Columns ascending.
Patterns build_patterns(const Mesh &m, const DofMap &d, bool expand_a00) {
  (void)m;
  Patterns P;
  const int dim = d.dim, NN = d.nn(), nv = dim + 1;
  std::vector<int64_t> nptr, vptr;
  std::vector<uint32_t> ninc, vinc;
  incidence(d.n_nodes, d.cell_nodes, NN, nptr, ninc);
  incidence(d.n_pverts, d.cell_pverts, nv, vptr, vinc);
  P.nodes = adjacency(d.n_nodes, d.n_nodes, nptr, ninc, d.cell_nodes, NN);
  const Csr n2p = adjacency(d.n_nodes, d.n_pverts, nptr, ninc, d.cell_pverts, nv);
  const Csr p2n = adjacency(d.n_pverts, d.n_nodes, vptr, vinc, d.cell_nodes, NN);
  // A01: row dim*A+c has the pressure columns of node A
  P.a01.n_rows = d.n_u;
  P.a01.n_cols = d.n_p;
  P.a01.rowptr.resize((size_t)d.n_u + 1);
  P.a01.colind.resize((size_t)n2p.nnz() * dim);
  for (uint32_t A = 0; A < d.n_nodes; ++A) {
    const int64_t b = n2p.rowptr[A], len = n2p.rowptr[A + 1] - b;
    for (int c = 0; c < dim; ++c) {
      const int64_t o = dim * b + c * len;
      P.a01.rowptr[(size_t)dim * A + c] = o;
      std::copy(n2p.colind.begin() + b, n2p.colind.begin() + b + len, P.a01.colind.begin() + o);
    }
  }
  P.a01.rowptr[d.n_u] = n2p.nnz() * dim;
  // A10: row V has columns dim*B+k for the nodes B around V
  P.a10.n_rows = d.n_p;
  P.a10.n_cols = d.n_u;
  P.a10.rowptr.resize((size_t)d.n_p + 1);
  P.a10.colind.resize((size_t)p2n.nnz() * dim);
  for (uint32_t V = 0; V <= d.n_pverts; ++V) P.a10.rowptr[V] = p2n.rowptr[V] * dim;
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < p2n.nnz(); ++k)
    for (int c = 0; c < dim; ++c) P.a10.colind[(size_t)k * dim + c] = dim * p2n.colind[k] + c;
  // S = A10 * A01 pattern: pressure vertices W such that some node is adjacent to both
  {
    P.s.n_rows = P.s.n_cols = d.n_p;
    P.s.rowptr.assign((size_t)d.n_p + 1, 0);
    std::vector<std::vector<uint32_t>> rows(d.n_p);
#pragma omp parallel
    {
      std::vector<uint32_t> buf;
#pragma omp for schedule(dynamic, 256)
      for (int64_t V = 0; V < (int64_t)d.n_p; ++V) {
        buf.clear();
        for (int64_t k = p2n.rowptr[V]; k < p2n.rowptr[V + 1]; ++k) {
          const uint32_t A = p2n.colind[k];
          buf.insert(buf.end(), n2p.colind.begin() + n2p.rowptr[A], n2p.colind.begin() + n2p.rowptr[A + 1]);
        }
        std::sort(buf.begin(), buf.end());
        buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
        rows[V] = buf;
      }
    }
    for (uint32_t V = 0; V < d.n_p; ++V) P.s.rowptr[V + 1] = P.s.rowptr[V] + (int64_t)rows[V].size();
    P.s.colind.resize((size_t)P.s.rowptr[d.n_p]);
#pragma omp parallel for schedule(static)
    for (int64_t V = 0; V < (int64_t)d.n_p; ++V)
      std::copy(rows[V].begin(), rows[V].end(), P.s.colind.begin() + P.s.rowptr[V]);
  }
  if (expand_a00) {
    P.a00.n_rows = P.a00.n_cols = d.n_u;
    P.a00.rowptr.resize((size_t)d.n_u + 1);
    P.a00.colind.resize((size_t)P.nodes.nnz() * dim * dim);
    for (uint32_t A = 0; A < d.n_nodes; ++A) {
      const int64_t b = P.nodes.rowptr[A], len = P.nodes.rowptr[A + 1] - b;
      for (int c = 0; c < dim; ++c) P.a00.rowptr[(size_t)dim * A + c] = dim * dim * b + c * dim * len;
    }
    P.a00.rowptr[d.n_u] = P.nodes.nnz() * dim * dim;
#pragma omp parallel for schedule(static)
    for (int64_t A = 0; A < (int64_t)d.n_nodes; ++A) {
      const int64_t b = P.nodes.rowptr[A], len = P.nodes.rowptr[A + 1] - b;
      for (int c = 0; c < dim; ++c) {
        uint32_t *o = &P.a00.colind[(size_t)(dim * dim * b + c * dim * len)];
        for (int64_t k = 0; k < len; ++k)
          for (int e = 0; e < dim; ++e) *o++ = dim * P.nodes.colind[b + k] + e;
      }
    }
  }
  return P;
}

std::vector<BoundaryFace> boundary_faces(const Mesh &m) {
  const int dim = m.dim, nv = dim + 1, nf = dim + 1;
  struct Rec {
    std::array<uint32_t, 3> key;
    uint32_t cell;
    int lf;
  };
  auto face_key = [&](const uint32_t *v, int f) {
    std::array<uint32_t, 3> k{0, 0, 0};
    if (dim == 2) {
      k[0] = v[kTriLines[f][0]];
      k[1] = v[kTriLines[f][1]];
      k[2] = UINT32_MAX;
    } else
      for (int r = 0; r < 3; ++r) k[r] = v[kTetFaces[f][r]];
    std::sort(k.begin(), k.end());
    return k;
  };
  std::vector<Rec> recs;
  recs.reserve(m.n_cells() * nf);
  for (size_t c = 0; c < m.n_cells(); ++c)
    for (int f = 0; f < nf; ++f) recs.push_back({face_key(&m.cells[c * nv], f), (uint32_t)c, f});
  std::sort(recs.begin(), recs.end(), [](const Rec &a, const Rec &b) {
    return a.key != b.key ? a.key < b.key : a.cell < b.cell;
  });
  std::map<std::array<uint32_t, 3>, int> tagged;
  for (size_t b = 0; b < m.n_bfaces(); ++b) {
    std::array<uint32_t, 3> k{0, 0, UINT32_MAX};
    for (int r = 0; r < dim; ++r) k[r] = m.bfaces[b * dim + r];
    std::sort(k.begin(), k.end());
    tagged[k] = m.bids[b];
  }
  std::vector<BoundaryFace> out;
  for (size_t i = 0; i < recs.size();) {
    size_t j = i + 1;
    while (j < recs.size() && recs[j].key == recs[i].key) ++j;
    if (j - i == 1) {
      auto it = tagged.find(recs[i].key);
      out.push_back({recs[i].cell, recs[i].lf, it == tagged.end() ? 0 : it->second});
    } else if (j - i > 2)
      throw std::runtime_error("boundary_faces: non-manifold mesh (a facet has more than two cells)");
    i = j;
  }
  std::sort(out.begin(), out.end(), [](const BoundaryFace &a, const BoundaryFace &b) {
    return a.cell != b.cell ? a.cell < b.cell : a.local_face < b.local_face;
  });
  return out;
}

DirichletSet dirichlet_dofs(const Mesh &m, const DofMap &d, const std::vector<BoundaryFace> &bf,
                            const std::function<double(const double *, int)> &profile) {
  const int dim = m.dim, nv = dim + 1;
  std::map<uint32_t, double> bv;
  auto visit = [&](const BoundaryFace &f, bool zero) {
    const uint32_t *v = &m.cells[(size_t)f.cell * nv];
    uint32_t fv[3];
    const int nfv = dim;
    for (int r = 0; r < nfv; ++r) fv[r] = dim == 2 ? v[kTriLines[f.local_face][r]] : v[kTetFaces[f.local_face][r]];
    uint32_t nodes[6];
    int nn = 0;
    for (int r = 0; r < nfv; ++r) nodes[nn++] = d.vert_node[fv[r]];
    if (dim == 2)
      nodes[nn++] = d.edge_node(fv[0], fv[1]);
    else
      for (int r = 0; r < 3; ++r) nodes[nn++] = d.edge_node(fv[r], fv[(r + 1) % 3]);
    for (int k = 0; k < nn; ++k)
      for (int c = 0; c < dim; ++c)
        bv[dim * nodes[k] + c] = zero ? 0.0 : profile(&d.node_xyz[(size_t)nodes[k] * dim], c);
  };
  for (auto &f : bf)
    if (f.id == 3) visit(f, false);
  // second interpolate_boundary_values call: std::map iteration order of the
  // function map is by boundary id (0, 2, 4) but the values are written while
  // walking the cells, so a dof shared by several ids gets the value of the
  // last face visited; all candidates agree for the reference's inlets.
  for (auto &f : bf)
    if (f.id == 0 || f.id == 2 || f.id == 4) visit(f, f.id == 4);
  DirichletSet s;
  for (auto &kv : bv) {
    s.dofs.push_back(kv.first);
    s.values.push_back(kv.second);
  }
  return s;
}

ForceFaces force_faces(const Mesh &m, const std::vector<BoundaryFace> &bf, int id) {
  ForceFaces F;
  const int dim = m.dim, nv = dim + 1;
  for (auto &f : bf) {
    if (f.id != id) continue;
    const uint32_t *v = &m.cells[(size_t)f.cell * nv];
    double n[3] = {0, 0, 0}, meas;
    const double *p[3];
    int opp_local = 0;
    bool on[4] = {false, false, false, false};
    for (int r = 0; r < dim; ++r) {
      const int lv = dim == 2 ? kTriLines[f.local_face][r] : kTetFaces[f.local_face][r];
      on[lv] = true;
      p[r] = &m.xyz[(size_t)v[lv] * dim];
    }
    for (int a = 0; a < nv; ++a)
      if (!on[a]) opp_local = a;
    const double *q = &m.xyz[(size_t)v[opp_local] * dim];
    if (dim == 2) {
      const double tx = p[1][0] - p[0][0], ty = p[1][1] - p[0][1];
      meas = std::hypot(tx, ty);
      n[0] = ty / meas;
      n[1] = -tx / meas;
    } else {
      const double a[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
      const double b[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
      n[0] = a[1] * b[2] - a[2] * b[1];
      n[1] = a[2] * b[0] - a[0] * b[2];
      n[2] = a[0] * b[1] - a[1] * b[0];
      const double l = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      meas = 0.5 * l;
      for (int r = 0; r < 3; ++r) n[r] /= l;
    }
    double s = 0;  // outward: pointing away from the opposite vertex
    for (int r = 0; r < dim; ++r) s += n[r] * (q[r] - p[0][r]);
    if (s > 0)
      for (int r = 0; r < dim; ++r) n[r] = -n[r];
    F.cell.push_back(f.cell);
    for (int r = 0; r < dim; ++r) F.normal.push_back(n[r]);
    F.measure.push_back(meas);
  }
  return F;
}

}  // namespace nsb
