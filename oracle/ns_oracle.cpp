// CPU oracle: a deliberately naive restatement of the reference's per-time-step
// path.  TEST INFRASTRUCTURE ONLY -- see ns_oracle.h (parity unpinned).
//
// Every routine cites the reference lines it follows (reference =
// /root/reference/src/NavierStokes.cpp unless another file is named) or the
// SURVEY.md appendix that restates the deal.II / Trilinos semantics.
#include "ns_oracle.h"
#include "ns_oracle_internal.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <vector>

namespace {

void pattern_from_rows(CsrMat &A, std::vector<std::vector<uint32_t>> &rows, int64_t n_cols) {
  A.n_rows = (int64_t)rows.size();
  A.n_cols = n_cols;
  A.rowptr.assign(rows.size() + 1, 0);
  for (size_t r = 0; r < rows.size(); ++r) {
    std::sort(rows[r].begin(), rows[r].end());
    rows[r].erase(std::unique(rows[r].begin(), rows[r].end()), rows[r].end());
    A.rowptr[r + 1] = A.rowptr[r] + (int64_t)rows[r].size();
  }
  A.colind.resize(A.rowptr.back());
  for (size_t r = 0; r < rows.size(); ++r) std::copy(rows[r].begin(), rows[r].end(), A.colind.begin() + A.rowptr[r]);
  A.val.assign(A.colind.size(), 0.0);
}

double dot(const Vec &a, const Vec &b) {
  double s = 0;
  for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
  return s;
}
double l2(const Vec &a) { return std::sqrt(dot(a, a)); }

// --------------------------------------------------------------------------
// ILU(0): TrilinosWrappers::PreconditionILU defaults (ilu_fill 0, atol 0,
// rtol 1, overlap 0) = Ifpack ILU(0) on the pattern of A.
// --------------------------------------------------------------------------
struct Ilu0 {
  const CsrMat *A = nullptr;
  std::vector<double> lu;
  std::vector<int64_t> diag;
  void initialize(const CsrMat &M) {
    A = &M;
    const int64_t n = M.n_rows;
    lu = M.val;
    diag.assign(n, -1);
    for (int64_t i = 0; i < n; ++i) diag[i] = M.find(i, (uint32_t)i);
    std::vector<int64_t> pos(M.n_cols, -1);
    for (int64_t i = 0; i < n; ++i) {
      for (int64_t k = M.rowptr[i]; k < M.rowptr[i + 1]; ++k) pos[M.colind[k]] = k;
      for (int64_t k = M.rowptr[i]; k < M.rowptr[i + 1] && M.colind[k] < (uint32_t)i; ++k) {
        const int64_t j = M.colind[k];
        const double l = lu[k] / lu[diag[j]];
        lu[k] = l;
        if (l != 0.0)
          for (int64_t kk = diag[j] + 1; kk < M.rowptr[j + 1]; ++kk) {
            const int64_t p = pos[M.colind[kk]];
            if (p >= 0) lu[p] -= l * lu[kk];
          }
      }
      for (int64_t k = M.rowptr[i]; k < M.rowptr[i + 1]; ++k) pos[M.colind[k]] = -1;
    }
  }
  void vmult(Vec &dst, const Vec &src) const {
    const int64_t n = A->n_rows;
    for (int64_t i = 0; i < n; ++i) {
      double s = src[i];
      for (int64_t k = A->rowptr[i]; k < diag[i]; ++k) s -= lu[k] * dst[A->colind[k]];
      dst[i] = s;
    }
    for (int64_t i = n - 1; i >= 0; --i) {
      double s = dst[i];
      for (int64_t k = diag[i] + 1; k < A->rowptr[i + 1]; ++k) s -= lu[k] * dst[A->colind[k]];
      dst[i] = s / lu[diag[i]];
    }
  }
};

// --------------------------------------------------------------------------
// deal.II SolverGMRES with default AdditionalData:
// max_n_tmp_vectors = 30 (restart length 28), left preconditioning, default
// (preconditioned) residual, modified Gram-Schmidt with the Kelley
// re-orthogonalisation test every fifth step.
// --------------------------------------------------------------------------
struct Gmres {
  int n_tmp = 30;
  int max_it = 10000;
  double tol = 0;  // absolute
  int last_step = 0;
  bool failed = false;
  std::vector<Vec> tmp;  // Krylov vectors live across restarts (stale contents: SURVEY.md B5)
  Vec p;

  template <class MatVec, class Prec>
  void solve(size_t n, const MatVec &A, Vec &x, const Vec &b, const Prec &P) {
    const int m = n_tmp - 2;
    if ((int)tmp.size() != n_tmp - 1) tmp.assign(n_tmp - 1, Vec());
    p.assign(n, 0.0);
    std::vector<double> H((size_t)(m + 1) * m, 0.0), gamma(m + 1), ci(m), si(m), h(m + 1);
    int acc = 0;
    bool iterate = true;
    failed = false;
    bool reorth = false;
    auto check = [&](int step, double res) {
      last_step = step;
      if (res <= tol) return false;  // success
      if (step >= max_it) {
        failed = true;
        return false;
      }
      return true;
    };
    auto tv = [&](int i) -> Vec & {
      if (tmp[i].size() != n) tmp[i].assign(n, 0.0);
      return tmp[i];
    };
    do {
      std::fill(H.begin(), H.end(), 0.0);
      Vec &v = tv(0);
      A(p, x);
      for (size_t i = 0; i < n; ++i) p[i] = b[i] - p[i];
      P(v, p);
      double rho = l2(v);
      iterate = check(acc, rho);
      if (!iterate) break;
      gamma[0] = rho;
      for (size_t i = 0; i < n; ++i) v[i] /= rho;
      int dim = 0;
      for (int it = 0; it < m && iterate; ++it) {
        ++acc;
        Vec &vv = tv(it + 1);
        A(p, tmp[it]);
        P(vv, p);
        dim = it + 1;
        // modified Gram-Schmidt
        const bool consider = !reorth && (it % 5 == 4);
        double norm_start = 0;
        if (consider) norm_start = l2(vv);
        for (int i = 0; i < dim; ++i) {
          h[i] = dot(vv, tmp[i]);
          for (size_t k = 0; k < n; ++k) vv[k] -= h[i] * tmp[i][k];
        }
        double s = l2(vv);
        if (consider && !(s > 10.0 * norm_start * std::sqrt(2.220446049250313e-16))) reorth = true;
        if (reorth) {
          for (int i = 0; i < dim; ++i) {
            const double t = dot(vv, tmp[i]);
            h[i] += t;
            for (size_t k = 0; k < n; ++k) vv[k] -= t * tmp[i][k];
          }
          s = l2(vv);
        }
        h[it + 1] = s;
        if (std::isfinite(1.0 / s))
          for (size_t k = 0; k < n; ++k) vv[k] /= s;
        // Givens rotations
        for (int i = 0; i < it; ++i) {
          const double d = h[i];
          h[i] = ci[i] * d + si[i] * h[i + 1];
          h[i + 1] = -si[i] * d + ci[i] * h[i + 1];
        }
        const double r = 1.0 / std::sqrt(h[it] * h[it] + h[it + 1] * h[it + 1]);
        si[it] = h[it + 1] * r;
        ci[it] = h[it] * r;
        h[it] = ci[it] * h[it] + si[it] * h[it + 1];
        gamma[it + 1] = -si[it] * gamma[it];
        gamma[it] *= ci[it];
        for (int i = 0; i < dim; ++i) H[(size_t)i * m + it] = h[i];
        rho = std::fabs(gamma[dim]);
        iterate = check(acc, rho);
      }
      // back substitution and update
      std::vector<double> y(dim);
      for (int i = dim - 1; i >= 0; --i) {
        double s = gamma[i];
        for (int j = i + 1; j < dim; ++j) s -= H[(size_t)i * m + j] * y[j];
        y[i] = s / H[(size_t)i * m + i];
      }
      for (int i = 0; i < dim; ++i)
        for (size_t k = 0; k < n; ++k) x[k] += y[i] * tmp[i][k];
    } while (iterate);
  }
};

// --------------------------------------------------------------------------
// Reference simplex tables
// --------------------------------------------------------------------------
const int TRI_LINES[3][2] = {{0, 1}, {1, 2}, {2, 0}};
const int TET_LINES[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
const int TET_FACES[4][3] = {{0, 1, 2}, {1, 0, 3}, {0, 2, 3}, {2, 1, 3}};

Quad cell_quadrature(int dim, int rule) {
  Quad q;
  if (dim == 2) {
    const double s15 = std::sqrt(15.0);
    const double p0 = (6 - s15) / 21, p3 = 1 - 2 * p0, p1 = (6 + s15) / 21, p2 = 1 - 2 * p1;
    const double w0 = 9.0 / 40, w1 = (155 - s15) / 1200, w2 = (155 + s15) / 1200;
    if (rule == NSO_QUAD_DEALII93) {
      // QGaussSimplex<2>(3) of deal.II 9.3.x: hard-coded 13-digit table
      q.pt = {{0.3333333333330, 0.3333333333330, 0}, {0.7974269853530, 0.1012865073230, 0},
              {0.1012865073230, 0.7974269853530, 0}, {0.1012865073230, 0.1012865073230, 0},
              {0.0597158717898, 0.4701420641050, 0}, {0.4701420641050, 0.0597158717898, 0},
              {0.4701420641050, 0.4701420641050, 0}};
      q.w = {0.5 * 0.225,          0.5 * 0.125939180545, 0.5 * 0.125939180545, 0.5 * 0.125939180545,
             0.5 * 0.132394152789, 0.5 * 0.132394152789, 0.5 * 0.132394152789};
    } else {
      // >= 9.4: QWitherdenVincentSimplex, barycentric permutations in
      // std::next_permutation order, first two coordinates are the point
      q.pt = {{1.0 / 3, 1.0 / 3, 0}, {p0, p0, 0}, {p0, p3, 0}, {p3, p0, 0}, {p2, p1, 0}, {p1, p2, 0}, {p1, p1, 0}};
      q.w = {0.5 * w0, 0.5 * w1, 0.5 * w1, 0.5 * w1, 0.5 * w2, 0.5 * w2, 0.5 * w2};
    }
  } else {
    if (rule == NSO_QUAD_DEALII93) {
      const double A = 0.5684305841968444, B = 0.1438564719343852;
      q.pt = {{A, B, B}, {B, B, B}, {B, B, A}, {B, A, B}, {0, .5, .5}, {.5, 0, .5}, {.5, .5, 0}, {.5, 0, 0}, {0, .5, 0}, {0, 0, .5}};
      for (int i = 0; i < 4; ++i) q.w.push_back(0.2177650698804054 / 6);
      for (int i = 0; i < 6; ++i) q.w.push_back(0.0214899534130631 / 6);
    } else {
      const double a1 = 3.1088591926330061e-01, b1 = 1 - 3 * a1, w1 = 1.1268792571801590e-01 / 6;
      const double a2 = 9.2735250310891248e-02, b2 = 1 - 3 * a2, w2 = 7.3493043116361956e-02 / 6;
      const double a3 = 4.5503704125649642e-02, b3 = (1 - 2 * a3) / 2, w3 = 4.2546020777081472e-02 / 6;
      q.pt = {{b1, a1, a1}, {a1, b1, a1}, {a1, a1, b1}, {a1, a1, a1},   // sorted (b1,a1,a1,a1) permutations
              {a2, a2, a2}, {a2, a2, b2}, {a2, b2, a2}, {b2, a2, a2},   // sorted (a2,a2,a2,b2)
              {a3, a3, b3}, {a3, b3, a3}, {a3, b3, b3}, {b3, a3, a3}, {b3, a3, b3}, {b3, b3, a3}};
      for (int i = 0; i < 4; ++i) q.w.push_back(w1);
      for (int i = 0; i < 4; ++i) q.w.push_back(w2);
      for (int i = 0; i < 6; ++i) q.w.push_back(w3);
    }
  }
  return q;
}

// weights of QGaussSimplex<dim-1>(3) normalised to sum 1 (times the face
// measure = JxW of FEFaceValues).
std::vector<double> face_weights(int dim, int rule) {
  if (dim == 2) return {5.0 / 18, 8.0 / 18, 5.0 / 18};
  const Quad q = cell_quadrature(2, rule);
  std::vector<double> w;
  for (double x : q.w) w.push_back(2.0 * x);
  return w;
}

// FE_SimplexP(2) / FE_SimplexP(1) on the reference simplex via barycentric
// coordinates; a < dim+1: vertex functions, then the line functions.
void p2_eval(int dim, const double *x, double *phi, double (*dphi)[3]) {
  const int nv = dim + 1;
  double lam[4], dl[4][3] = {{0}};
  lam[0] = 1;
  for (int d = 0; d < dim; ++d) {
    lam[0] -= x[d];
    lam[d + 1] = x[d];
    dl[0][d] = -1;
    dl[d + 1][d] = 1;
  }
  for (int a = 0; a < nv; ++a) {
    phi[a] = lam[a] * (2 * lam[a] - 1);
    for (int d = 0; d < dim; ++d) dphi[a][d] = (4 * lam[a] - 1) * dl[a][d];
  }
  const int nl = dim == 2 ? 3 : 6;
  for (int l = 0; l < nl; ++l) {
    const int *e = dim == 2 ? TRI_LINES[l] : TET_LINES[l];
    phi[nv + l] = 4 * lam[e[0]] * lam[e[1]];
    for (int d = 0; d < dim; ++d) dphi[nv + l][d] = 4 * (lam[e[0]] * dl[e[1]][d] + lam[e[1]] * dl[e[0]][d]);
  }
}
void p1_eval(int dim, const double *x, double *psi) {
  psi[0] = 1;
  for (int d = 0; d < dim; ++d) {
    psi[0] -= x[d];
    psi[d + 1] = x[d];
  }
}

}  // namespace

// ==========================================================================
// InletVelocity::value of the drivers (tests/2D/test_01/src/test_01.cpp:29-36,
// tests/3D/test_01/src/test_01.cpp:29-36, tests/2D/test_naca/src/test_03.cpp:28-35,
// tests/2D/test_03/src/test_03.cpp for the sin(pi t/8) factor).
double nso::inlet_value(const double *p, int comp, double t) const {
  if (comp != 0) return 0.0;
  double v;
  if (inlet_kind == NSO_INLET_UNIFORM)
    v = U_m;
  else if (dim == 2)
    v = 4 * U_m * p[1] * (H - p[1]) / (H * H);
  else
    v = 16 * U_m * p[1] * p[2] * (H - p[1]) * (H - p[2]) / (H * H * H * H);
  if (inlet_sin) v *= std::sin(M_PI * t / 8.0);
  return v;
}
double nso::mean_vel(double t) const {
  double v = inlet_kind == NSO_INLET_UNIFORM ? U_m : (dim == 2 ? 2.0 * U_m / 3.0 : 4.0 * U_m / 9.0);
  if (inlet_sin) v *= std::sin(M_PI * t / 8.0);
  return v;
}

extern "C" {

nso *nso_create(int dim, int64_t n_verts, const double *xyz, int64_t n_cells, const uint32_t *cells,
                int64_t n_bfaces, const uint32_t *bfaces, const int32_t *bids, int quad_rule) {
  nso *o = new nso;
  o->dim = dim;
  o->nv = dim + 1;
  o->nl = dim == 2 ? 3 : 6;
  o->NN = o->nv + o->nl;
  o->dpc = dim * o->NN + o->nv;
  o->rule = quad_rule;
  o->n_verts = n_verts;
  o->n_cells = n_cells;
  o->xyz.assign(xyz, xyz + n_verts * dim);
  o->cells.assign(cells, cells + n_cells * (dim + 1));
  o->bfaces.assign(bfaces, bfaces + n_bfaces * dim);
  o->bids.assign(bids, bids + n_bfaces);
  const int nv = o->nv, nl = o->nl, dpc = o->dpc;

  // ---- distribute_dofs + component_wise (reference :65-70) ----
  // pass 1: deal.II's cell walk hands out dim+1 consecutive indices per new
  // vertex and dim per new line.
  std::vector<int64_t> vert_first(n_verts, -1);
  std::map<std::pair<uint32_t, uint32_t>, int64_t> line_first;
  std::vector<int> comp_of;  // component of every old index
  std::vector<int64_t> old_cell_dofs((size_t)n_cells * dpc);
  for (int64_t c = 0; c < n_cells; ++c) {
    const uint32_t *v = &o->cells[c * nv];
    for (int a = 0; a < nv; ++a)
      if (vert_first[v[a]] < 0) {
        vert_first[v[a]] = (int64_t)comp_of.size();
        for (int k = 0; k <= dim; ++k) comp_of.push_back(k);
      }
    for (int l = 0; l < nl; ++l) {
      const int *e = dim == 2 ? TRI_LINES[l] : TET_LINES[l];
      auto key = std::minmax(v[e[0]], v[e[1]]);
      if (!line_first.count(key)) {
        line_first[key] = (int64_t)comp_of.size();
        for (int k = 0; k < dim; ++k) comp_of.push_back(k);
      }
    }
    int64_t *od = &old_cell_dofs[(size_t)c * dpc];
    for (int a = 0; a < nv; ++a)
      for (int k = 0; k <= dim; ++k) *od++ = vert_first[v[a]] + k;
    for (int l = 0; l < nl; ++l) {
      const int *e = dim == 2 ? TRI_LINES[l] : TET_LINES[l];
      const int64_t f = line_first[std::minmax(v[e[0]], v[e[1]])];
      for (int k = 0; k < dim; ++k) *od++ = f + k;
    }
  }
  // pass 2: stable partition into blocks {velocity, pressure}
  const size_t N = comp_of.size();
  std::vector<uint32_t> renum(N);
  uint32_t nu_ = 0, np_ = 0;
  for (size_t i = 0; i < N; ++i)
    if (comp_of[i] < dim) ++nu_;
  o->n_u = nu_;
  {
    uint32_t iu = 0;
    for (size_t i = 0; i < N; ++i)
      if (comp_of[i] < dim)
        renum[i] = iu++;
      else
        renum[i] = nu_ + np_++;
  }
  o->n_p = np_;
  o->cell_dofs.resize((size_t)n_cells * dpc);
  for (size_t i = 0; i < o->cell_dofs.size(); ++i) o->cell_dofs[i] = renum[old_cell_dofs[i]];
  // local dof -> (component, scalar shape index)
  o->local_comp.resize(dpc);
  o->local_scalar.resize(dpc);
  {
    int i = 0;
    for (int a = 0; a < nv; ++a)
      for (int k = 0; k <= dim; ++k, ++i) {
        o->local_comp[i] = k;
        o->local_scalar[i] = a;
      }
    for (int l = 0; l < nl; ++l)
      for (int k = 0; k < dim; ++k, ++i) {
        o->local_comp[i] = k;
        o->local_scalar[i] = nv + l;
      }
  }
  // support points (vertices; line midpoints under the affine mapping)
  o->support.assign(N, {0, 0, 0});
  for (int64_t c = 0; c < n_cells; ++c) {
    const uint32_t *v = &o->cells[c * nv];
    for (int i = 0; i < dpc; ++i) {
      const int a = o->local_scalar[i];
      std::array<double, 3> p{0, 0, 0};
      if (a < nv)
        for (int r = 0; r < dim; ++r) p[r] = o->xyz[(size_t)v[a] * dim + r];
      else {
        const int *e = dim == 2 ? TRI_LINES[a - nv] : TET_LINES[a - nv];
        for (int r = 0; r < dim; ++r)
          p[r] = 0.5 * (o->xyz[(size_t)v[e[0]] * dim + r] + o->xyz[(size_t)v[e[1]] * dim + r]);
      }
      o->support[o->cell_dofs[(size_t)c * dpc + i]] = p;
    }
  }

  // ---- sparsity (reference :101-117) ----
  {
    std::vector<std::vector<uint32_t>> r00(nu_), r01(nu_), r10(np_);
    for (int64_t c = 0; c < n_cells; ++c) {
      const uint32_t *d = &o->cell_dofs[(size_t)c * dpc];
      for (int i = 0; i < dpc; ++i)
        for (int j = 0; j < dpc; ++j) {
          const bool pi = o->local_comp[i] == dim, pj = o->local_comp[j] == dim;
          if (pi && pj) continue;  // coupling[dim][dim] = none
          if (!pi && !pj) r00[d[i]].push_back(d[j]);
          if (!pi && pj) r01[d[i]].push_back(d[j] - nu_);
          if (pi && !pj) r10[d[i] - nu_].push_back(d[j]);
        }
    }
    pattern_from_rows(o->A00, r00, nu_);
    pattern_from_rows(o->A01, r01, np_);
    pattern_from_rows(o->A10, r10, nu_);
    // S = B * diag * Bt structural product (reference :956)
    std::vector<std::vector<uint32_t>> rs(np_);
    for (uint32_t i = 0; i < np_; ++i)
      for (int64_t k = o->A10.rowptr[i]; k < o->A10.rowptr[i + 1]; ++k) {
        const uint32_t u = o->A10.colind[k];
        rs[i].insert(rs[i].end(), o->A01.colind.begin() + o->A01.rowptr[u], o->A01.colind.begin() + o->A01.rowptr[u + 1]);
      }
    pattern_from_rows(o->S, rs, np_);
  }
  // ---- boundary faces with ids (untagged boundary faces keep id 0) ----
  {
    std::map<std::array<uint32_t, 3>, std::vector<std::pair<uint32_t, int>>> faces;
    for (int64_t c = 0; c < n_cells; ++c)
      for (int f = 0; f < nv; ++f) {
        std::array<uint32_t, 3> k{0, 0, UINT32_MAX};
        for (int r = 0; r < dim; ++r)
          k[r] = o->cells[c * nv + (dim == 2 ? TRI_LINES[f][r] : TET_FACES[f][r])];
        std::sort(k.begin(), k.end());
        faces[k].push_back({(uint32_t)c, f});
      }
    std::map<std::array<uint32_t, 3>, int> tag;
    for (int64_t b = 0; b < n_bfaces; ++b) {
      std::array<uint32_t, 3> k{0, 0, UINT32_MAX};
      for (int r = 0; r < dim; ++r) k[r] = bfaces[b * dim + r];
      std::sort(k.begin(), k.end());
      tag[k] = bids[b];
    }
    for (auto &kv : faces)
      if (kv.second.size() == 1) {
        auto it = tag.find(kv.first);
        o->bf.push_back({kv.second[0].first, kv.second[0].second, it == tag.end() ? 0 : it->second});
      }
    std::sort(o->bf.begin(), o->bf.end(), [](const nso::BFace &a, const nso::BFace &b) {
      return a.cell != b.cell ? a.cell < b.cell : a.lf < b.lf;
    });
  }
  // ---- reference tables (FEValues with update_values | update_gradients) ----
  o->quad = cell_quadrature(dim, quad_rule);
  o->wface = face_weights(dim, quad_rule);
  o->nq = (int)o->quad.w.size();
  o->nqf = (int)o->wface.size();
  o->phi.resize((size_t)o->nq * o->NN);
  o->dphi.resize((size_t)o->nq * o->NN * 3);
  o->psi.resize((size_t)o->nq * nv);
  for (int q = 0; q < o->nq; ++q) {
    double ph[10], dph[10][3], ps[4];
    p2_eval(dim, o->quad.pt[q].data(), ph, dph);
    p1_eval(dim, o->quad.pt[q].data(), ps);
    for (int a = 0; a < o->NN; ++a) {
      o->phi[(size_t)q * o->NN + a] = ph[a];
      for (int d = 0; d < 3; ++d) o->dphi[((size_t)q * o->NN + a) * 3 + d] = d < dim ? dph[a][d] : 0.0;
    }
    for (int a = 0; a < nv; ++a) o->psi[(size_t)q * nv + a] = ps[a];
  }
  const size_t Ntot = (size_t)nu_ + np_;
  o->rhs.assign(Ntot, 0.0);
  o->lumped.assign(Ntot, 0.0);
  o->solution_owned.assign(Ntot, 0.0);  // InitialConditions == 0 (NavierStokes.hpp:140-163)
  o->solution.assign(Ntot, 0.0);
  return o;
}

void nso_destroy(nso *o) {
  if (o) nso_baseline_free(o);
  delete o;
}

void nso_sizes(const nso *o, int64_t out[10]) {
  out[0] = o->n_u;
  out[1] = o->n_p;
  out[2] = o->A00.rowptr.back();
  out[3] = o->A01.rowptr.back();
  out[4] = o->A10.rowptr.back();
  out[5] = o->S.rowptr.back();
  out[6] = o->dpc;
  out[7] = o->nq;
  out[8] = o->nqf;
  out[9] = (int64_t)o->bc_dofs.size();
}
void nso_get_cell_dofs(const nso *o, uint32_t *out) { std::copy(o->cell_dofs.begin(), o->cell_dofs.end(), out); }
static const CsrMat &blk(const nso *o, int b) { return b == 0 ? o->A00 : b == 1 ? o->A01 : b == 2 ? o->A10 : o->S; }
void nso_get_pattern(const nso *o, int block, int64_t *rowptr, uint32_t *colind) {
  const CsrMat &A = blk(o, block);
  std::copy(A.rowptr.begin(), A.rowptr.end(), rowptr);
  std::copy(A.colind.begin(), A.colind.end(), colind);
}
void nso_get_values(const nso *o, int block, double *vals) {
  const CsrMat &A = blk(o, block);
  std::copy(A.val.begin(), A.val.end(), vals);
}
void nso_get_rhs(const nso *o, double *out) { std::copy(o->rhs.begin(), o->rhs.end(), out); }
void nso_get_lumped(const nso *o, double *out) { std::copy(o->lumped.begin(), o->lumped.end(), out); }
void nso_get_bc(const nso *o, uint32_t *dofs, double *values) {
  std::copy(o->bc_dofs.begin(), o->bc_dofs.end(), dofs);
  std::copy(o->bc_vals.begin(), o->bc_vals.end(), values);
}
void nso_set_params(nso *o, double deltat, double nu) {
  o->deltat = deltat;
  o->nu = nu;
}
void nso_set_bc_diag_mode(nso *o, int mode) { o->bc_diag_mode = mode; }
void nso_set_inlet(nso *o, int kind, double U_m, double H, int time_sin) {
  o->inlet_kind = kind;
  o->U_m = U_m;
  o->H = H;
  o->inlet_sin = time_sin;
}
double nso_mean_velocity(const nso *o, double time) { return o->mean_vel(time); }
double nso_set_re_number(nso *o, int Re) {  // reference :332-341
  const double U = o->mean_vel(o->inlet_time);
  o->nu = (U * o->Diameter) / Re;
  return o->nu;
}
void nso_set_solution(nso *o, const double *x) {
  o->solution_owned.assign(x, x + o->solution_owned.size());
  o->solution = o->solution_owned;
}
void nso_get_solution(const nso *o, double *x) { std::copy(o->solution_owned.begin(), o->solution_owned.end(), x); }
void nso_set_solver(nso *o, double outer_rtol, int n_tmp_vectors, int max_it, double inner_rtol) {
  o->outer_rtol = outer_rtol;
  o->n_tmp = n_tmp_vectors;
  o->max_it = max_it;
  o->inner_rtol = inner_rtol;
}
void nso_set_threads(nso *o, int n) { o->threads = n < 1 ? 1 : n; }

// --------------------------------------------------------------------------
// FEValues::reinit for the affine MappingFE: physical
// gradients J^{-T} grad_hat, JxW = |det J| w_q.
// --------------------------------------------------------------------------
static void cell_geometry(const nso *o, int64_t c, double Jinv[3][3], double *absdet) {
  const int dim = o->dim;
  const uint32_t *v = &o->cells[c * o->nv];
  double J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int a = 0; a < dim; ++a)
    for (int r = 0; r < dim; ++r) J[r][a] = o->xyz[(size_t)v[a + 1] * dim + r] - o->xyz[(size_t)v[0] * dim + r];
  double det;
  if (dim == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    Jinv[0][0] = J[1][1] / det;
    Jinv[0][1] = -J[0][1] / det;
    Jinv[1][0] = -J[1][0] / det;
    Jinv[1][1] = J[0][0] / det;
  } else {
    det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
          J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    Jinv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
    Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
    Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
    Jinv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
    Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
    Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    Jinv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
    Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
    Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  }
  *absdet = std::fabs(det);
}

static void reinit_cell(const nso *o, int64_t c, CellFE &fe) {
  const int dim = o->dim, dpc = o->dpc, nq = o->nq, NN = o->NN, nv = o->nv;
  double Jinv[3][3], adet;
  cell_geometry(o, c, Jinv, &adet);
  fe.val.assign((size_t)nq * dpc * 3, 0.0);
  fe.grad.assign((size_t)nq * dpc * 9, 0.0);
  fe.div.assign((size_t)nq * dpc, 0.0);
  fe.pval.assign((size_t)nq * dpc, 0.0);
  fe.JxW.resize(nq);
  for (int q = 0; q < nq; ++q) {
    fe.JxW[q] = adet * o->quad.w[q];
    for (int i = 0; i < dpc; ++i) {
      const int comp = o->local_comp[i], a = o->local_scalar[i];
      if (comp == dim) {
        fe.pval[(size_t)q * dpc + i] = o->psi[(size_t)q * nv + a];
        continue;
      }
      fe.val[((size_t)q * dpc + i) * 3 + comp] = o->phi[(size_t)q * NN + a];
      double g[3] = {0, 0, 0};  // J^{-T} grad_hat
      for (int e = 0; e < dim; ++e)
        for (int d = 0; d < dim; ++d) g[e] += Jinv[d][e] * o->dphi[((size_t)q * NN + a) * 3 + d];
      for (int e = 0; e < dim; ++e) fe.grad[((size_t)q * dpc + i) * 9 + comp * 3 + e] = g[e];
      fe.div[(size_t)q * dpc + i] = g[comp];
    }
  }
}

// reference :171-254: the (q, i, j) triple loop, literally.
void nso_cell_contribution(const nso *o, int64_t c, CellFE &fe, double *cell_matrix, double *cell_rhs,
                           double *cell_lumped) {
  const int dim = o->dim, dpc = o->dpc, nq = o->nq;
  reinit_cell(o, c, fe);
  std::fill(cell_matrix, cell_matrix + dpc * dpc, 0.0);
  std::fill(cell_rhs, cell_rhs + dpc, 0.0);
  std::fill(cell_lumped, cell_lumped + dpc, 0.0);
  const uint32_t *dofs = &o->cell_dofs[(size_t)c * dpc];
  for (int q = 0; q < nq; ++q) {
    // :175 get_function_values of the ghosted solution
    double u[3] = {0, 0, 0};
    for (int i = 0; i < dpc; ++i)
      for (int d = 0; d < dim; ++d) u[d] += o->solution[dofs[i]] * fe.val[((size_t)q * dpc + i) * 3 + d];
    const double f[3] = {0, 0, 0};  // ForcingTerm == 0 (NavierStokes.hpp:56-65)
    const double JxW = fe.JxW[q];
    for (int i = 0; i < dpc; ++i) {
      const double *vi = &fe.val[((size_t)q * dpc + i) * 3];
      const double *gi = &fe.grad[((size_t)q * dpc + i) * 9];
      double temp = 0;
      for (int j = 0; j < dpc; ++j) {
        const double *vj = &fe.val[((size_t)q * dpc + j) * 3];
        const double *gj = &fe.grad[((size_t)q * dpc + j) * 9];
        double m = 0, k = 0, t1 = 0;
        for (int d = 0; d < dim; ++d) m += vi[d] * vj[d];  // :191-194
        for (int d = 0; d < dim; ++d)
          for (int e = 0; e < dim; ++e) k += gi[d * 3 + e] * gj[d * 3 + e];  // :197-200
        for (int e = 0; e < dim; ++e) {                                      // :204-208 value(i)*gradient(j)*u
          double w = 0;
          for (int d = 0; d < dim; ++d) w += vi[d] * gj[d * 3 + e];
          t1 += w * u[e];
        }
        double a = cell_matrix[i * dpc + j];
        a += m * JxW / o->deltat;
        a += o->nu * k * JxW;
        a += t1 * JxW;
        a -= fe.div[(size_t)q * dpc + i] * fe.pval[(size_t)q * dpc + j] * JxW;  // :222-224
        a -= fe.div[(size_t)q * dpc + j] * fe.pval[(size_t)q * dpc + i] * JxW;  // :227-229
        cell_matrix[i * dpc + j] = a;
        temp += std::fabs(m * JxW);  // :232-236
      }
      double fv = 0, uv = 0;
      for (int d = 0; d < dim; ++d) {
        fv += f[d] * vi[d];
        uv += u[d] * vi[d];
      }
      cell_rhs[i] += fv * JxW;               // :241-243
      cell_rhs[i] += uv * JxW / o->deltat;   // :245-248
      cell_lumped[i] += temp;                // :252
    }
  }
  // :257-278 Neumann term: -p_out * int n.v over faces with id 1.  p_out is a
  // const 0.0 (NavierStokes.hpp:255), so the face loop contributes exactly 0.
  (void)o->p_out;
}

void nso_assemble(nso *o, double time) {
  const int dpc = o->dpc;
  const uint32_t nu_ = o->n_u;
  std::fill(o->A00.val.begin(), o->A00.val.end(), 0.0);  // :154-156
  std::fill(o->A01.val.begin(), o->A01.val.end(), 0.0);
  std::fill(o->A10.val.begin(), o->A10.val.end(), 0.0);
  std::fill(o->rhs.begin(), o->rhs.end(), 0.0);
  std::fill(o->lumped.begin(), o->lumped.end(), 0.0);
  // element matrices are computed in batches (optionally by several threads);
  // the scatter below is serial and in cell order, so sums keep the
  // reference's order.
  const int64_t batch = 256;
  std::vector<double> M((size_t)batch * dpc * dpc), R((size_t)batch * dpc), L((size_t)batch * dpc);
  for (int64_t c0 = 0; c0 < o->n_cells; c0 += batch) {
    const int64_t nb = std::min(batch, o->n_cells - c0);
#pragma omp parallel num_threads(o->threads)
    {
      CellFE fe;
#pragma omp for schedule(static)
      for (int64_t b = 0; b < nb; ++b)
        nso_cell_contribution(o, c0 + b, fe, &M[(size_t)b * dpc * dpc], &R[(size_t)b * dpc], &L[(size_t)b * dpc]);
    }
    for (int64_t b = 0; b < nb; ++b) {
      const uint32_t *dofs = &o->cell_dofs[(size_t)(c0 + b) * dpc];
      const double *cm = &M[(size_t)b * dpc * dpc];
      // :282 BlockSparseMatrix::add elides zero values (the (p,p) entries and the
      // cross-component velocity couplings are exactly 0.0)
      for (int i = 0; i < dpc; ++i)
        for (int j = 0; j < dpc; ++j) {
          const double v = cm[i * dpc + j];
          if (v == 0.0) continue;
          const bool pi = dofs[i] >= nu_, pj = dofs[j] >= nu_;
          CsrMat &A = !pi ? (!pj ? o->A00 : o->A01) : o->A10;
          A.val[A.find(pi ? dofs[i] - nu_ : dofs[i], pj ? dofs[j] - nu_ : dofs[j])] += v;
        }
      for (int i = 0; i < dpc; ++i) {
        o->rhs[dofs[i]] += R[(size_t)b * dpc + i];     // :283
        o->lumped[dofs[i]] += L[(size_t)b * dpc + i];  // :284
      }
    }
  }
  for (auto &x : o->lumped) x = o->deltat / x;  // :287-290 (deltat/0 on pressure dofs)

  nso_apply_boundary(o, time);
}

// ---- Dirichlet boundary conditions, :297-329 ----
void nso_apply_boundary(nso *o, double time) {
  const int dim = o->dim, dpc = o->dpc;
  const uint32_t nu_ = o->n_u;
  o->inlet_time = time;  // :306 inlet_velocity.set_time(time)
  std::map<uint32_t, double> bv;
  auto interpolate = [&](const std::vector<int> &ids, const std::vector<bool> &zero) {
    for (const auto &f : o->bf) {
      int which = -1;
      for (size_t k = 0; k < ids.size(); ++k)
        if (ids[k] == f.id) which = (int)k;
      if (which < 0) continue;
      // dofs on the face: those of the face's vertices and lines
      bool on_face[4] = {false, false, false, false};
      for (int r = 0; r < dim; ++r) on_face[dim == 2 ? TRI_LINES[f.lf][r] : TET_FACES[f.lf][r]] = true;
      for (int i = 0; i < dpc; ++i) {
        const int comp = o->local_comp[i], a = o->local_scalar[i];
        if (comp == dim) continue;  // ComponentMask: velocity only (:300-301)
        bool on;
        if (a < o->nv)
          on = on_face[a];
        else {
          const int *e = dim == 2 ? TRI_LINES[a - o->nv] : TET_LINES[a - o->nv];
          on = on_face[e[0]] && on_face[e[1]];
        }
        if (!on) continue;
        const uint32_t dof = o->cell_dofs[(size_t)f.cell * dpc + i];
        bv[dof] = zero[which] ? 0.0 : o->inlet_value(o->support[dof].data(), comp, time);
      }
    }
  };
  interpolate({3}, {false});                      // :307-311
  interpolate({0, 2, 4}, {false, false, true});   // :313-324
  o->bc_dofs.clear();
  o->bc_vals.clear();
  for (auto &kv : bv) {
    o->bc_dofs.push_back(kv.first);
    o->bc_vals.push_back(kv.second);
  }
  // MatrixTools::apply_boundary_values(..., eliminate_columns = false), :326-328
  // (SURVEY.md A.7): first non-zero diagonal entry of the block; constrained
  // rows of A00 cleared keeping a non-zero diagonal (clear_row), same rows of
  // A01 cleared; rhs = diag * g; solution = g.
  if (!bv.empty()) {
    double first_diag = 1.0;
    for (uint32_t i = 0; i < nu_; ++i) {
      const double dgn = o->A00.val[o->A00.find(i, i)];
      if (dgn != 0.0) {
        first_diag = std::fabs(dgn);
        break;
      }
    }
    for (auto &kv : bv) {
      const uint32_t i = kv.first;
      const int64_t dpos = o->A00.find(i, i);
      for (int64_t k = o->A00.rowptr[i]; k < o->A00.rowptr[i + 1]; ++k)
        if (k != dpos) o->A00.val[k] = 0.0;
      if (o->bc_diag_mode == 1 || o->A00.val[dpos] == 0.0) o->A00.val[dpos] = first_diag;
      for (int64_t k = o->A01.rowptr[i]; k < o->A01.rowptr[i + 1]; ++k) o->A01.val[k] = 0.0;
      o->solution_owned[i] = kv.second;
      o->rhs[i] = kv.second * o->A00.val[dpos];
    }
  }
}

void nso_vmult(const nso *o, const double *x, double *y) {
  const uint32_t nu_ = o->n_u, np_ = o->n_p;
  Vec t(nu_);
  o->A00.vmult(y, x);
  o->A01.vmult(t.data(), x + nu_);
  for (uint32_t i = 0; i < nu_; ++i) y[i] += t[i];
  o->A10.vmult(y + nu_, x);
  (void)np_;
}

// reference :344-397 + PreconditionASIMPLE :934-995
int nso_solve_time_step(nso *o, int *iters, double *t_prec, double *t_solve) {
  const uint32_t nu_ = o->n_u, np_ = o->n_p;
  const size_t N = (size_t)nu_ + np_;
  auto t0 = std::chrono::high_resolution_clock::now();
  const double tol = o->outer_rtol * l2(o->rhs);  // :348
  // --- PreconditionASIMPLE::initialize, :934-963 ---
  Vec Di(nu_);
  for (uint32_t i = 0; i < nu_; ++i) Di[i] = 1.0 / o->A00.val[o->A00.find(i, i)];  // :948-953
  {  // :956  S = B * diag(Di) * Bt
    std::fill(o->S.val.begin(), o->S.val.end(), 0.0);
    for (uint32_t i = 0; i < np_; ++i)
      for (int64_t k = o->A10.rowptr[i]; k < o->A10.rowptr[i + 1]; ++k) {
        const uint32_t u = o->A10.colind[k];
        const double bu = o->A10.val[k] * Di[u];
        for (int64_t kk = o->A01.rowptr[u]; kk < o->A01.rowptr[u + 1]; ++kk)
          o->S.val[o->S.find(i, o->A01.colind[kk])] += bu * o->A01.val[kk];
      }
  }
  Ilu0 precF, precS;
  precF.initialize(o->A00);  // :958
  precS.initialize(o->S);    // :959
  Vec vec0(nu_, 0.0), vec1(np_, 0.0);  // :961-962 (reinit zeroes)
  auto t1 = std::chrono::high_resolution_clock::now();

  Gmres innerF, innerS;
  innerF.n_tmp = innerS.n_tmp = 30;
  innerF.max_it = innerS.max_it = 10000;  // :972
  bool inner_failed = false;
  Vec src0(nu_), src1(np_), d0(nu_), d1(np_);
  // --- PreconditionASIMPLE::vmult, :966-995 ---
  auto Pvmult = [&](Vec &dst, const Vec &src) {
    std::copy(src.begin(), src.begin() + nu_, src0.begin());
    std::copy(src.begin() + nu_, src.end(), src1.begin());
    std::copy(dst.begin() + nu_, dst.end(), d1.begin());  // stale initial guess
    innerF.tol = o->inner_rtol * l2(src0);                 // :978
    innerF.solve(nu_, [&](Vec &y, const Vec &x) { o->A00.vmult(y.data(), x.data()); }, vec0, src0,
                 [&](Vec &y, const Vec &x) { precF.vmult(y, x); });  // :981
    o->A10.vmult(vec1.data(), vec0.data());                            // :982
    for (uint32_t i = 0; i < np_; ++i) vec1[i] = -vec1[i] + src1[i];   // :983 sadd(-1, src1)
    innerS.tol = o->inner_rtol * l2(vec1);                             // :986
    innerS.solve(np_, [&](Vec &y, const Vec &x) { o->S.vmult(y.data(), x.data()); }, d1, vec1,
                 [&](Vec &y, const Vec &x) { precS.vmult(y, x); });  // :989
    inner_failed = inner_failed || innerF.failed || innerS.failed;
    for (auto &x : d1) x *= -1.0 / o->alpha;                           // :990
    o->A01.vmult(d0.data(), d1.data());                                // :992
    for (uint32_t i = 0; i < nu_; ++i) d0[i] = -(d0[i] * Di[i]) + vec0[i];  // :993-994 scale(Di); sadd(-1, vec0)
    std::copy(d0.begin(), d0.end(), dst.begin());
    std::copy(d1.begin(), d1.end(), dst.begin() + nu_);
  };
  Gmres outer;
  outer.n_tmp = o->n_tmp;
  outer.max_it = o->max_it;
  outer.tol = tol;
  outer.solve(N, [&](Vec &y, const Vec &x) { nso_vmult(o, x.data(), y.data()); }, o->solution_owned, o->rhs, Pvmult);  // :377
  auto t2 = std::chrono::high_resolution_clock::now();
  if (iters) *iters = outer.last_step;
  if (t_prec) *t_prec = std::chrono::duration<double>(t1 - t0).count();
  if (t_solve) *t_solve = std::chrono::duration<double>(t2 - t1).count();
  o->solution = o->solution_owned;  // :395
  return (outer.failed || inner_failed) ? 1 : 0;
}

// reference :831-929, including quirks B2-B4 of SURVEY.md: the cell-interior
// quadrature values are indexed with the face quadrature counter.
void nso_compute_forces(nso *o, double time, double out[4]) {
  (void)time;
  const int dim = o->dim, dpc = o->dpc, nq = o->nq, nqf = o->nqf, nv = o->nv;
  double ldrag = 0, llift = 0;
  CellFE fe;
  for (const auto &f : o->bf) {
    if (f.id != 4) continue;  // :874-875
    const int64_t c = f.cell;
    reinit_cell(o, c, fe);  // :864
    const uint32_t *dofs = &o->cell_dofs[(size_t)c * dpc];
    // :866-868 values at the CELL quadrature points
    std::vector<double> pq(nq, 0.0), gq((size_t)nq * 9, 0.0);
    for (int q = 0; q < nq; ++q)
      for (int i = 0; i < dpc; ++i) {
        const double s = o->solution[dofs[i]];
        pq[q] += s * fe.pval[(size_t)q * dpc + i];
        for (int k = 0; k < 9; ++k) gq[(size_t)q * 9 + k] += s * fe.grad[((size_t)q * dpc + i) * 9 + k];
      }
    // :877 FEFaceValues::reinit: outward normal and JxW = w_q * measure
    const uint32_t *v = &o->cells[c * nv];
    const double *p[3];
    bool on[4] = {false, false, false, false};
    for (int r = 0; r < dim; ++r) {
      const int lv = dim == 2 ? TRI_LINES[f.lf][r] : TET_FACES[f.lf][r];
      on[lv] = true;
      p[r] = &o->xyz[(size_t)v[lv] * dim];
    }
    int opp = 0;
    for (int a = 0; a < nv; ++a)
      if (!on[a]) opp = a;
    const double *qo = &o->xyz[(size_t)v[opp] * dim];
    double n[3] = {0, 0, 0}, meas;
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
    double s = 0;
    for (int r = 0; r < dim; ++r) s += n[r] * (qo[r] - p[0][r]);
    if (s > 0)
      for (int r = 0; r < dim; ++r) n[r] = -n[r];
    for (int q = 0; q < nqf; ++q) {  // :879-903
      const double nx = n[0], ny = n[1];
      const double tangent[3] = {ny, -nx, 0.0};
      const double JxW = o->wface[q] * meas;
      double ngt = 0;  // normal * grad(u) * tangent
      for (int e = 0; e < dim; ++e) {
        double w = 0;
        for (int d = 0; d < dim; ++d) w += n[d] * gq[(size_t)q * 9 + d * 3 + e];
        ngt += w * tangent[e];
      }
      ldrag += o->nu * ngt * ny * JxW;
      ldrag -= pq[q] * nx * JxW;
      llift -= o->nu * ngt * nx * JxW;
      llift -= pq[q] * ny * JxW;
    }
  }
  const double U = o->mean_vel(o->inlet_time);  // :911 (inlet time = last set_time in assemble)
  double cd, cl;
  if (dim == 3) {  // :913-922
    cd = 2.0 * -ldrag / (U * U * o->Diameter * 0.41);
    cl = 2.0 * -llift / (U * U * o->Diameter * 0.41);
  } else {
    cd = 2.0 * -ldrag / (U * U * o->Diameter);
    cl = 2.0 * -llift / (U * U * o->Diameter);
  }
  out[0] = ldrag;
  out[1] = llift;
  out[2] = cd;
  out[3] = cl;
}

}  // extern "C"
