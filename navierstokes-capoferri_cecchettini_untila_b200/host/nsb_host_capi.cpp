// C ABI of the host setup library (include/nsb_host.h).
#include "../../include/nsb_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>

#include "distribute.hpp"
#include "problem.hpp"

struct nsh_problem {
  nsb::Problem p;
};
struct nsh_local {
  nsb::LocalProblem l;
};

namespace {
thread_local std::string g_err;
template <class F>
int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  } catch (...) {
    g_err = "unknown error";
    return -2;
  }
}
}  // namespace

namespace nsb {
// Recursive coordinate bisection on cell centroids: split the longest axis of
// the current box at the weighted median so that part sizes differ by at most
// one cell; parts are numbered left to right.
void Problem::partition(int n_parts) {
  const int dim = mesh.dim, nv = dim + 1;
  const size_t nc = mesh.n_cells();
  std::vector<double> cen(nc * dim, 0.0);
  for (size_t c = 0; c < nc; ++c)
    for (int a = 0; a < nv; ++a)
      for (int r = 0; r < dim; ++r) cen[c * dim + r] += mesh.xyz[(size_t)mesh.cells[c * nv + a] * dim + r] / nv;
  part_cell.assign(nc, 0);
  std::vector<uint32_t> idx(nc);
  std::iota(idx.begin(), idx.end(), 0u);
  struct Job {
    size_t b, e;
    int p0, np;
  };
  std::vector<Job> stack{{0, nc, 0, n_parts}};
  while (!stack.empty()) {
    Job j = stack.back();
    stack.pop_back();
    if (j.np == 1) {
      for (size_t k = j.b; k < j.e; ++k) part_cell[idx[k]] = j.p0;
      continue;
    }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (size_t k = j.b; k < j.e; ++k)
      for (int r = 0; r < dim; ++r) {
        lo[r] = std::min(lo[r], cen[(size_t)idx[k] * dim + r]);
        hi[r] = std::max(hi[r], cen[(size_t)idx[k] * dim + r]);
      }
    int ax = 0;
    for (int r = 1; r < dim; ++r)
      if (hi[r] - lo[r] > hi[ax] - lo[ax]) ax = r;
    const int npl = j.np / 2;
    const size_t mid = j.b + (j.e - j.b) * npl / j.np;
    std::nth_element(idx.begin() + j.b, idx.begin() + mid, idx.begin() + j.e, [&](uint32_t a, uint32_t b) {
      const double xa = cen[(size_t)a * dim + ax], xb = cen[(size_t)b * dim + ax];
      return xa != xb ? xa < xb : a < b;
    });
    stack.push_back({j.b, mid, j.p0, npl});
    stack.push_back({mid, j.e, j.p0 + npl, j.np - npl});
  }
}
}  // namespace nsb

extern "C" {

const char *nsh_last_error(void) { return g_err.c_str(); }

int nsh_problem_generate(const char *name, double h, nsh_problem **out) {
  return guarded([&] {
    auto *P = new nsh_problem;
    try {
      P->p.mesh = nsb::gen_named(name, h);
    } catch (...) {
      delete P;
      throw;
    }
    *out = P;
  });
}
int nsh_problem_generate_airfoil(const char *dat_path, int naca4, double chord, double aoa_deg, double Lx, double Ly,
                                 double cx, double cy, double h, nsh_problem **out) {
  return guarded([&] {
    if (!(chord > 0) || !(h > 0) || !out) throw std::runtime_error("nsh_problem_generate_airfoil: bad arguments");
    const int n_around = std::max(32, (int)std::lround(2.1 * chord / h));
    const auto unit = (dat_path && *dat_path) ? nsb::read_airfoil_dat(dat_path) : nsb::naca4_contour(naca4, n_around);
    auto *P = new nsh_problem;
    try {
      P->p.mesh = nsb::gen_airfoil2d(Lx, Ly, cx, cy, nsb::place_airfoil(unit, chord, aoa_deg, cx, cy), chord, 48);
    } catch (...) {
      delete P;
      throw;
    }
    *out = P;
  });
}
int nsh_problem_read(const char *msh_path, int dim, nsh_problem **out) {
  return guarded([&] {
    auto *P = new nsh_problem;
    try {
      P->p.mesh = nsb::read_msh(msh_path, dim);
    } catch (...) {
      delete P;
      throw;
    }
    *out = P;
  });
}
int nsh_problem_from_arrays(int dim, int64_t n_verts, const double *xyz, int64_t n_cells, const uint32_t *cells,
                            int64_t n_bfaces, const uint32_t *bfaces, const int32_t *bids, nsh_problem **out) {
  return guarded([&] {
    if (dim != 2 && dim != 3) throw std::runtime_error("dim must be 2 or 3");
    auto *P = new nsh_problem;
    nsb::Mesh &m = P->p.mesh;
    m.dim = dim;
    m.xyz.assign(xyz, xyz + n_verts * dim);
    m.cells.assign(cells, cells + n_cells * (dim + 1));
    m.bfaces.assign(bfaces, bfaces + n_bfaces * dim);
    m.bids.assign(bids, bids + n_bfaces);
    nsb::orient_cells(m);
    *out = P;
  });
}
int nsh_problem_write_msh(const nsh_problem *P, const char *path) {
  return guarded([&] { nsb::write_msh(P->p.mesh, path); });
}
void nsh_problem_free(nsh_problem *P) { delete P; }

int nsh_build_space(nsh_problem *P, int expand_a00) {
  return guarded([&] { P->p.build_space(expand_a00 != 0); });
}
int nsh_set_inlet(nsh_problem *P, int kind, double U_m, double H, int time_sin) {
  P->p.inlet.kind = kind;
  P->p.inlet.U_m = U_m;
  P->p.inlet.H = H;
  P->p.inlet.time_sin = time_sin;
  return 0;
}
int nsh_build_boundary(nsh_problem *P) {
  return guarded([&] {
    if (!P->p.has_space) throw std::runtime_error("nsh_build_boundary: call nsh_build_space first");
    P->p.build_boundary();
  });
}
double nsh_mean_velocity(const nsh_problem *P, double time) { return P->p.inlet.mean_vel(P->p.mesh.dim, time); }
double nsh_inlet_time_factor(const nsh_problem *P, double time) { return P->p.inlet.time_factor(time); }

int nsh_sizes(const nsh_problem *P, int64_t out[10]) {
  const nsb::Problem &p = P->p;
  out[0] = p.mesh.dim;
  out[1] = (int64_t)p.mesh.n_verts();
  out[2] = (int64_t)p.mesh.n_cells();
  out[3] = (int64_t)p.mesh.n_bfaces();
  out[4] = p.dofs.n_nodes;
  out[5] = p.dofs.n_u;
  out[6] = p.dofs.n_p;
  out[7] = p.has_space ? p.dofs.dofs_per_cell() : 0;
  out[8] = (int64_t)p.bc.dofs.size();
  out[9] = (int64_t)p.ff.cell.size();
  return 0;
}

int nsh_array(const nsh_problem *P, const char *name, const void **data, int64_t *count, int *elem_bytes) {
  const nsb::Problem &p = P->p;
  const std::string n(name);
  auto set = [&](const auto &v) {
    *data = v.data();
    *count = (int64_t)v.size();
    *elem_bytes = (int)sizeof(v[0]);
    return 0;
  };
  if (n == "xyz") return set(p.mesh.xyz);
  if (n == "cells") return set(p.mesh.cells);
  if (n == "bfaces") return set(p.mesh.bfaces);
  if (n == "bids") return set(p.mesh.bids);
  if (n == "cell_dofs") return set(p.dofs.cell_dofs);
  if (n == "cell_nodes") return set(p.dofs.cell_nodes);
  if (n == "cell_pverts") return set(p.dofs.cell_pverts);
  if (n == "node_xyz") return set(p.dofs.node_xyz);
  if (n == "bc.dofs") return set(p.bc.dofs);
  if (n == "bc.values") return set(p.bc.values);
  if (n == "ff.cell") return set(p.ff.cell);
  if (n == "ff.normal") return set(p.ff.normal);
  if (n == "ff.measure") return set(p.ff.measure);
  if (n == "part.cell") return set(p.part_cell);
  const size_t dot = n.find('.');
  if (dot != std::string::npos) {
    const std::string b = n.substr(0, dot), f = n.substr(dot + 1);
    const nsb::Csr *A = b == "nodes" ? &p.pat.nodes
                        : b == "a00" ? &p.pat.a00
                        : b == "a01" ? &p.pat.a01
                        : b == "a10" ? &p.pat.a10
                        : b == "s"   ? &p.pat.s
                                     : nullptr;
    if (A && f == "rowptr") return set(A->rowptr);
    if (A && f == "colind") return set(A->colind);
  }
  g_err = "nsh_array: unknown array '" + n + "'";
  return -1;
}

int nsh_partition(nsh_problem *P, int n_parts) {
  return guarded([&] {
    if (n_parts < 1) throw std::runtime_error("nsh_partition: n_parts < 1");
    P->p.partition(n_parts);
  });
}

int nsh_localize(const nsh_problem *P, int n_parts, int rank, nsh_local **out) {
  return guarded([&] {
    auto *L = new nsh_local;
    try {
      L->l = nsb::localize(P->p, n_parts, rank);
    } catch (...) {
      delete L;
      throw;
    }
    *out = L;
  });
}
void nsh_local_free(nsh_local *L) { delete L; }
int nsh_local_sizes(const nsh_local *L, int64_t out[11]) {
  const nsb::LocalProblem &l = L->l;
  out[0] = l.n_own;
  out[1] = l.n_ghost;
  out[2] = l.n_p;
  out[3] = l.n_p_own();
  out[4] = l.p_offset[l.rank];
  out[5] = (int64_t)l.cells.size();
  out[6] = (int64_t)l.neighbors.size();
  out[7] = (int64_t)l.bc_nodes.size();
  out[8] = (int64_t)l.ff_cell.size();
  out[9] = l.n_nodes_global;
  out[10] = l.node_offset[l.rank];
  return 0;
}
int nsh_local_array(const nsh_local *L, const char *name, const void **data, int64_t *count, int *elem_bytes) {
  const nsb::LocalProblem &l = L->l;
  const std::string n(name);
  auto set = [&](const auto &v) {
    *data = v.data();
    *count = (int64_t)v.size();
    *elem_bytes = (int)sizeof(v[0]);
    return 0;
  };
  if (n == "node_offset") return set(l.node_offset);
  if (n == "p_offset") return set(l.p_offset);
  if (n == "node_perm") return set(l.node_perm);
  if (n == "p_perm") return set(l.p_perm);
  if (n == "ghost_dist") return set(l.ghost_dist);
  if (n == "cells") return set(l.cells);
  if (n == "cell_verts") return set(l.cell_verts);
  if (n == "cell_nodes") return set(l.cell_nodes);
  if (n == "cell_pverts") return set(l.cell_pverts);
  if (n == "neighbors") return set(l.neighbors);
  if (n == "send_ptr") return set(l.send_ptr);
  if (n == "recv_ptr") return set(l.recv_ptr);
  if (n == "send_idx") return set(l.send_idx);
  if (n == "bc_nodes") return set(l.bc_nodes);
  if (n == "bc_values") return set(l.bc_values);
  if (n == "ff.cell") return set(l.ff_cell);
  if (n == "ff.normal") return set(l.ff_normal);
  if (n == "ff.measure") return set(l.ff_measure);
  const size_t dot = n.find('.');
  if (dot != std::string::npos) {
    const std::string b = n.substr(0, dot), f = n.substr(dot + 1);
    const nsb::Csr *A = b == "fs" ? &l.fs : b == "a01" ? &l.a01 : b == "a10" ? &l.a10 : b == "s" ? &l.s : nullptr;
    if (A && f == "rowptr") return set(A->rowptr);
    if (A && f == "colind") return set(A->colind);
  }
  g_err = "nsh_local_array: unknown array '" + n + "'";
  return -1;
}

}  // extern "C"
