// See distribute.hpp.
#include "distribute.hpp"

#include <algorithm>
#include <stdexcept>

namespace nsb {

namespace {
// rows of `src` listed in `rows`, columns mapped by `colmap` and re-sorted
template <class ColMap>
Csr extract_rows(const Csr &src, const std::vector<uint32_t> &rows, uint32_t n_cols, ColMap colmap) {
  Csr A;
  A.n_rows = (uint32_t)rows.size();
  A.n_cols = n_cols;
  A.rowptr.assign(rows.size() + 1, 0);
  for (size_t r = 0; r < rows.size(); ++r) A.rowptr[r + 1] = A.rowptr[r] + (src.rowptr[rows[r] + 1] - src.rowptr[rows[r]]);
  A.colind.resize((size_t)A.rowptr.back());
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < (int64_t)rows.size(); ++r) {
    uint32_t *o = &A.colind[(size_t)A.rowptr[r]];
    const int64_t b = src.rowptr[rows[r]], e = src.rowptr[rows[r] + 1];
    for (int64_t k = b; k < e; ++k) o[k - b] = colmap(src.colind[k]);
    std::sort(o, o + (e - b));
  }
  return A;
}
}  // namespace

LocalProblem localize(const Problem &P, int n_parts, int rank) {
  if (!P.has_space || !P.has_boundary) throw std::runtime_error("localize: build the space and boundary lists first");
  if ((int)P.part_cell.size() != (int)P.mesh.n_cells()) throw std::runtime_error("localize: partition the problem first");
  if (rank < 0 || rank >= n_parts) throw std::runtime_error("localize: bad rank");
  const DofMap &d = P.dofs;
  const int dim = d.dim, NN = d.nn(), nv = dim + 1;
  const size_t nc = P.mesh.n_cells();
  LocalProblem L;
  L.rank = rank;
  L.n_parts = n_parts;
  L.dim = dim;
  L.n_nodes_global = d.n_nodes;
  L.n_p = d.n_p;
  // ownership: lowest part touching the entity
  std::vector<int32_t> own_n(d.n_nodes, INT32_MAX), own_p(d.n_pverts, INT32_MAX);
  for (size_t c = 0; c < nc; ++c) {
    const int32_t part = P.part_cell[c];
    if (part < 0 || part >= n_parts) throw std::runtime_error("localize: partition has more parts than n_parts");
    for (int a = 0; a < NN; ++a) own_n[d.cell_nodes[c * NN + a]] = std::min(own_n[d.cell_nodes[c * NN + a]], part);
    for (int a = 0; a < nv; ++a) own_p[d.cell_pverts[c * nv + a]] = std::min(own_p[d.cell_pverts[c * nv + a]], part);
  }
  auto make_perm = [&](const std::vector<int32_t> &own, std::vector<uint32_t> &off, std::vector<uint32_t> &perm) {
    off.assign(n_parts + 1, 0);
    for (int32_t o : own) ++off[o + 1];
    for (int q = 0; q < n_parts; ++q) off[q + 1] += off[q];
    std::vector<uint32_t> fill(off.begin(), off.end() - 1);
    perm.resize(own.size());
    for (size_t i = 0; i < own.size(); ++i) perm[i] = fill[own[i]]++;
  };
  make_perm(own_n, L.node_offset, L.node_perm);
  make_perm(own_p, L.p_offset, L.p_perm);
  L.n_own = L.node_offset[rank + 1] - L.node_offset[rank];
  // local cells: touch an owned node
  for (size_t c = 0; c < nc; ++c) {
    bool mine = P.part_cell[c] == rank;
    for (int a = 0; a < NN && !mine; ++a) mine = own_n[d.cell_nodes[c * NN + a]] == rank;
    if (mine) L.cells.push_back((uint32_t)c);
  }
  // ghosts
  for (uint32_t c : L.cells)
    for (int a = 0; a < NN; ++a) {
      const uint32_t A = d.cell_nodes[(size_t)c * NN + a];
      if (own_n[A] != rank) L.ghost_dist.push_back(L.node_perm[A]);
    }
  std::sort(L.ghost_dist.begin(), L.ghost_dist.end());
  L.ghost_dist.erase(std::unique(L.ghost_dist.begin(), L.ghost_dist.end()), L.ghost_dist.end());
  L.n_ghost = (uint32_t)L.ghost_dist.size();
  const uint32_t off_r = L.node_offset[rank];
  auto local_of = [&](uint32_t A) -> uint32_t {  // canonical node -> local id
    const uint32_t g = L.node_perm[A];
    if (own_n[A] == rank) return g - off_r;
    auto it = std::lower_bound(L.ghost_dist.begin(), L.ghost_dist.end(), g);
    if (it == L.ghost_dist.end() || *it != g) return UINT32_MAX;
    return L.n_own + (uint32_t)(it - L.ghost_dist.begin());
  };
  L.cell_verts.resize(L.cells.size() * nv);
  L.cell_nodes.resize(L.cells.size() * NN);
  L.cell_pverts.resize(L.cells.size() * nv);
  for (size_t i = 0; i < L.cells.size(); ++i) {
    const size_t c = L.cells[i];
    for (int a = 0; a < nv; ++a) {
      L.cell_verts[i * nv + a] = P.mesh.cells[c * nv + a];
      L.cell_pverts[i * nv + a] = L.p_perm[d.cell_pverts[c * nv + a]];
    }
    for (int a = 0; a < NN; ++a) L.cell_nodes[i * NN + a] = local_of(d.cell_nodes[c * NN + a]);
  }
  // owned entities in distributed (= canonical relative) order
  std::vector<uint32_t> own_nodes, own_pv;
  for (uint32_t A = 0; A < d.n_nodes; ++A)
    if (own_n[A] == rank) own_nodes.push_back(A);
  for (uint32_t V = 0; V < d.n_pverts; ++V)
    if (own_p[V] == rank) own_pv.push_back(V);
  // local patterns
  L.fs = extract_rows(P.pat.nodes, own_nodes, L.n_own + L.n_ghost, [&](uint32_t B) { return local_of(B); });
  {
    std::vector<uint32_t> rows;
    rows.reserve(own_nodes.size() * dim);
    for (uint32_t A : own_nodes)
      for (int c = 0; c < dim; ++c) rows.push_back(dim * A + c);
    L.a01 = extract_rows(P.pat.a01, rows, d.n_p, [&](uint32_t V) { return L.p_perm[V]; });
  }
  L.a10 = extract_rows(P.pat.a10, own_pv, dim * (L.n_own + L.n_ghost), [&](uint32_t u) {
    const uint32_t a = local_of(u / dim);  // an unmapped node must not wrap around in dim * a + c
    return a == UINT32_MAX ? UINT32_MAX : (uint32_t)(dim * a + u % dim);
  });
  {
    std::vector<uint32_t> inv(d.n_pverts);
    for (uint32_t V = 0; V < d.n_pverts; ++V) inv[L.p_perm[V]] = V;
    L.s = extract_rows(P.pat.s, inv, d.n_p, [&](uint32_t W) { return L.p_perm[W]; });
  }
  for (uint32_t v : L.fs.colind)
    if (v == UINT32_MAX) throw std::runtime_error("localize: a column of an owned row is not local");
  for (uint32_t v : L.a10.colind)
    if (v >= (uint32_t)dim * (L.n_own + L.n_ghost))
      throw std::runtime_error("localize: a velocity column of an owned pressure row is not local");
  // halo lists: node B owned by r is a ghost on q iff a cell containing B also contains a node owned by q
  {
    std::vector<std::pair<int32_t, uint32_t>> send, recv;  // (peer, distributed node id)
    for (size_t c = 0; c < nc; ++c) {
      const uint32_t *cn = &d.cell_nodes[c * NN];
      // ranks that hold this cell: the owners of its nodes and the part that owns the cell
      int32_t holders[11];
      int nh = 0;
      auto add = [&](int32_t q) {
        for (int i = 0; i < nh; ++i)
          if (holders[i] == q) return;
        holders[nh++] = q;
      };
      add(P.part_cell[c]);
      for (int a = 0; a < NN; ++a) add(own_n[cn[a]]);
      if (nh == 1) continue;
      for (int a = 0; a < NN; ++a) {
        const int32_t o = own_n[cn[a]];
        for (int i = 0; i < nh; ++i) {
          const int32_t q = holders[i];
          if (q == o) continue;
          if (o == rank) send.push_back({q, L.node_perm[cn[a]]});  // my node is a ghost on q
          if (q == rank) recv.push_back({o, L.node_perm[cn[a]]});  // o's node is my ghost
        }
      }
    }
    auto uniq = [](std::vector<std::pair<int32_t, uint32_t>> &v) {
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
    };
    uniq(send);
    uniq(recv);
    {  // neighbours: every peer in either list (the relation is symmetric for edge-connected partitions)
      std::vector<int32_t> nb;
      for (auto &s : send) nb.push_back(s.first);
      for (auto &r : recv) nb.push_back(r.first);
      std::sort(nb.begin(), nb.end());
      nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
      L.neighbors = nb;
    }
    L.send_ptr.assign(L.neighbors.size() + 1, 0);
    L.recv_ptr.assign(L.neighbors.size() + 1, 0);
    for (size_t k = 0; k < L.neighbors.size(); ++k) {
      L.send_ptr[k + 1] = L.send_ptr[k];
      L.recv_ptr[k + 1] = L.recv_ptr[k];
      for (auto &s : send)
        if (s.first == L.neighbors[k]) {
          L.send_idx.push_back(s.second - off_r);
          ++L.send_ptr[k + 1];
        }
      for (auto &r : recv)
        if (r.first == L.neighbors[k]) ++L.recv_ptr[k + 1];
    }
    // the ghosts are sorted by distributed id = grouped by owner in rank order, like `recv`
    if ((uint32_t)recv.size() != L.n_ghost) throw std::runtime_error("localize: ghost list and receive list differ");
    for (size_t i = 0; i < recv.size(); ++i)
      if (recv[i].second != L.ghost_dist[i]) throw std::runtime_error("localize: ghost order mismatch");
  }
  // Dirichlet nodes (node-complete list, canonical order) owned by this rank
  for (size_t i = 0; i + dim <= P.bc.dofs.size(); i += dim) {
    const uint32_t A = P.bc.dofs[i] / dim;
    if (P.bc.dofs[i] % dim != 0) throw std::runtime_error("localize: Dirichlet list is not node-complete");
    if (own_n[A] != rank) continue;
    L.bc_nodes.push_back(L.node_perm[A] - off_r);
    for (int c = 0; c < dim; ++c) L.bc_values.push_back(P.bc.values[i + c]);
  }
  // obstacle faces in the cells this rank owns (reference :861: is_locally_owned)
  for (size_t f = 0; f < P.ff.cell.size(); ++f) {
    const uint32_t c = P.ff.cell[f];
    if (P.part_cell[c] != rank) continue;
    auto it = std::lower_bound(L.cells.begin(), L.cells.end(), c);
    if (it == L.cells.end() || *it != c) throw std::runtime_error("localize: owned cell missing from the local list");
    L.ff_cell.push_back((uint32_t)(it - L.cells.begin()));
    for (int r = 0; r < dim; ++r) L.ff_normal.push_back(P.ff.normal[f * dim + r]);
    L.ff_measure.push_back(P.ff.measure[f]);
  }
  return L;
}

}  // namespace nsb
