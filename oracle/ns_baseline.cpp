// CPU BASELINE (bench.py only): the reference's per-time-step path the way
// `mpirun -n P` runs it on P cores, restated with P OpenMP threads in one address
// space.  TEST/BENCH INFRASTRUCTURE ONLY -- see ns_oracle.h (parity unpinned).
//
// What `mpirun -n P` of the reference does (reference = src/NavierStokes.cpp):
//   * :19-23   the cells are partitioned into P subdomains; a DoF belongs to the
//              lowest subdomain touching it (deal.II subdomain-wise numbering);
//   * :164-285 every rank assembles its own cells; off-rank rows travel in
//              compress(add) (:292-294) -- here: atomic adds into the shared CSR;
//   * :958-959 TrilinosWrappers::PreconditionILU = Ifpack ILU(0) with overlap 0:
//              each rank factorises only the diagonal block of ITS rows, couplings
//              to other ranks' DoFs are dropped (block-Jacobi ILU);
//   * :377, :979-989 every SpMV, dot product and vector update of the outer and the
//              two inner GMRES solves runs on all ranks over their own rows.
// The serial oracle (ns_oracle.cpp) is the checker and is not touched by this
// mode; the two share the element routine and the boundary-condition routine.
#include <omp.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ns_oracle_internal.h"

namespace {

int g_threads = 1;
// vectors shorter than this are handled by one thread: a fork/join costs more than the loop (an MPI rank
// would pay a few microseconds of all-reduce latency instead)
constexpr int64_t kParMin = 16384;

// Ifpack-style local ILU(0) of the diagonal blocks of a row partition.
struct LocalIlu {
  struct Part {
    std::vector<uint32_t> rows;      // global rows of this subdomain, ascending
    std::vector<int64_t> rowptr;     // local CSR of the diagonal block
    std::vector<uint32_t> col;       // local column ids
    std::vector<int64_t> src;        // position of every kept entry in the global matrix
    std::vector<int64_t> diag;
    std::vector<double> lu;
  };
  std::vector<Part> parts;
  void setup(const CsrMat &M, const std::vector<int32_t> &owner, int P) {
    parts.assign(P, Part());
    std::vector<uint32_t> g2l(M.n_rows);
    for (int64_t i = 0; i < M.n_rows; ++i) {
      Part &p = parts[owner[i]];
      g2l[i] = (uint32_t)p.rows.size();
      p.rows.push_back((uint32_t)i);
    }
#pragma omp parallel for schedule(dynamic, 1) num_threads(P)
    for (int q = 0; q < P; ++q) {
      Part &p = parts[q];
      p.rowptr.assign(p.rows.size() + 1, 0);
      p.diag.assign(p.rows.size(), -1);
      for (size_t li = 0; li < p.rows.size(); ++li) {
        const int64_t i = p.rows[li];
        for (int64_t k = M.rowptr[i]; k < M.rowptr[i + 1]; ++k) {
          const uint32_t j = M.colind[k];
          if (owner[j] != q) continue;  // coupling to another rank's DoF: dropped (overlap 0)
          if (j == (uint32_t)i) p.diag[li] = (int64_t)p.col.size();
          p.col.push_back(g2l[j]);
          p.src.push_back(k);
        }
        p.rowptr[li + 1] = (int64_t)p.col.size();
      }
      p.lu.resize(p.col.size());
    }
  }
  // numeric factorisation, all subdomains concurrently
  void factor(const CsrMat &M, int P) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(P)
    for (int q = 0; q < P; ++q) {
      Part &p = parts[q];
      const int64_t n = (int64_t)p.rows.size();
      for (size_t k = 0; k < p.src.size(); ++k) p.lu[k] = M.val[p.src[k]];
      std::vector<int64_t> pos(n, -1);
      for (int64_t i = 0; i < n; ++i) {
        for (int64_t k = p.rowptr[i]; k < p.rowptr[i + 1]; ++k) pos[p.col[k]] = k;
        for (int64_t k = p.rowptr[i]; k < p.rowptr[i + 1] && p.col[k] < (uint32_t)i; ++k) {
          const int64_t j = p.col[k];
          const double l = p.lu[k] / p.lu[p.diag[j]];
          p.lu[k] = l;
          if (l != 0.0)
            for (int64_t kk = p.diag[j] + 1; kk < p.rowptr[j + 1]; ++kk) {
              const int64_t t = pos[p.col[kk]];
              if (t >= 0) p.lu[t] -= l * p.lu[kk];
            }
        }
        for (int64_t k = p.rowptr[i]; k < p.rowptr[i + 1]; ++k) pos[p.col[k]] = -1;
      }
    }
  }
  void vmult(Vec &dst, const Vec &src, int P) const {
#pragma omp parallel for schedule(dynamic, 1) num_threads(P) if ((int64_t)src.size() > kParMin / 8)
    for (int q = 0; q < P; ++q) {
      const Part &p = parts[q];
      const int64_t n = (int64_t)p.rows.size();
      std::vector<double> y(n);
      for (int64_t i = 0; i < n; ++i) {
        double s = src[p.rows[i]];
        for (int64_t k = p.rowptr[i]; k < p.diag[i]; ++k) s -= p.lu[k] * y[p.col[k]];
        y[i] = s;
      }
      for (int64_t i = n - 1; i >= 0; --i) {
        double s = y[i];
        for (int64_t k = p.diag[i] + 1; k < p.rowptr[i + 1]; ++k) s -= p.lu[k] * y[p.col[k]];
        y[i] = s / p.lu[p.diag[i]];
      }
      for (int64_t i = 0; i < n; ++i) dst[p.rows[i]] = y[i];
    }
  }
};


double pdot(const Vec &a, const Vec &b) {
  double s = 0;
  const int64_t n = (int64_t)a.size();
#pragma omp parallel for reduction(+ : s) schedule(static) num_threads(g_threads) if (n > kParMin)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
double pl2(const Vec &a) { return std::sqrt(pdot(a, a)); }
void paxpy(Vec &y, double a, const Vec &x) {
  const int64_t n = (int64_t)y.size();
#pragma omp parallel for schedule(static) num_threads(g_threads) if (n > kParMin)
  for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}
void pscale(Vec &y, double a) {
  const int64_t n = (int64_t)y.size();
#pragma omp parallel for schedule(static) num_threads(g_threads) if (n > kParMin)
  for (int64_t i = 0; i < n; ++i) y[i] *= a;
}
void pvmult(const CsrMat &A, double *y, const double *x) {
#pragma omp parallel for schedule(static) num_threads(g_threads) if (A.n_rows > kParMin / 8)
  for (int64_t r = 0; r < A.n_rows; ++r) {
    double s = 0;
    for (int64_t k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k) s += A.val[k] * x[A.colind[k]];
    y[r] = s;
  }
}

// deal.II SolverGMRES as in ns_oracle.cpp, vector operations on all threads
struct ParGmres {
  int n_tmp = 30, max_it = 10000, last_step = 0;
  double tol = 0;
  bool failed = false;
  std::vector<Vec> tmp;  // stale across calls
  Vec p;
  template <class MatVec, class Prec>
  void solve(size_t n, const MatVec &A, Vec &x, const Vec &b, const Prec &P) {
    const int m = n_tmp - 2;
    if ((int)tmp.size() != n_tmp - 1) tmp.assign(n_tmp - 1, Vec());
    p.assign(n, 0.0);
    std::vector<double> H((size_t)(m + 1) * m, 0.0), gamma(m + 1), ci(m), si(m), h(m + 1);
    int acc = 0;
    bool iterate = true, reorth = false;
    failed = false;
    auto check = [&](int step, double res) {
      last_step = step;
      if (res <= tol) return false;
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
      const int64_t nn = (int64_t)n;
#pragma omp parallel for schedule(static) num_threads(g_threads) if (nn > kParMin)
      for (int64_t i = 0; i < nn; ++i) p[i] = b[i] - p[i];
      P(v, p);
      double rho = pl2(v);
      iterate = check(acc, rho);
      if (!iterate) break;
      gamma[0] = rho;
      pscale(v, 1.0 / rho);
      int dim = 0;
      for (int it = 0; it < m && iterate; ++it) {
        ++acc;
        Vec &vv = tv(it + 1);
        A(p, tmp[it]);
        P(vv, p);
        dim = it + 1;
        const bool consider = !reorth && (it % 5 == 4);
        double norm_start = 0;
        if (consider) norm_start = pl2(vv);
        for (int i = 0; i < dim; ++i) {
          h[i] = pdot(vv, tmp[i]);
          paxpy(vv, -h[i], tmp[i]);
        }
        double s = pl2(vv);
        if (consider && !(s > 10.0 * norm_start * std::sqrt(2.220446049250313e-16))) reorth = true;
        if (reorth) {
          for (int i = 0; i < dim; ++i) {
            const double t = pdot(vv, tmp[i]);
            h[i] += t;
            paxpy(vv, -t, tmp[i]);
          }
          s = pl2(vv);
        }
        h[it + 1] = s;
        if (std::isfinite(1.0 / s)) pscale(vv, 1.0 / s);
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
      std::vector<double> y(dim);
      for (int i = dim - 1; i >= 0; --i) {
        double s = gamma[i];
        for (int j = i + 1; j < dim; ++j) s -= H[(size_t)i * m + j] * y[j];
        y[i] = s / H[(size_t)i * m + i];
      }
      for (int i = 0; i < dim; ++i) paxpy(x, y[i], tmp[i]);
    } while (iterate);
  }
};

}  // namespace

struct NsoBaseline {
  int P = 1;
  std::vector<int32_t> cell_part, u_owner, p_owner;
  std::vector<std::vector<int64_t>> cells_of;
  LocalIlu iluF, iluS;
};

extern "C" {

void nso_baseline_free(nso *o) {
  delete o->baseline;
  o->baseline = nullptr;
}

int nso_baseline_partition(nso *o, int n_parts, const int32_t *cell_part) {
  if (n_parts < 1) return 1;
  nso_baseline_free(o);
  NsoBaseline *B = new NsoBaseline;
  o->baseline = B;
  B->P = n_parts;
  B->cell_part.assign(cell_part, cell_part + o->n_cells);
  B->cells_of.assign(n_parts, {});
  // DoF owner = lowest subdomain touching it (deal.II)
  B->u_owner.assign(o->n_u, INT32_MAX);
  B->p_owner.assign(o->n_p, INT32_MAX);
  for (int64_t c = 0; c < o->n_cells; ++c) {
    const int32_t q = cell_part[c];
    if (q < 0 || q >= n_parts) return 1;
    B->cells_of[q].push_back(c);
    const uint32_t *d = &o->cell_dofs[(size_t)c * o->dpc];
    for (int i = 0; i < o->dpc; ++i) {
      int32_t &w = d[i] < o->n_u ? B->u_owner[d[i]] : B->p_owner[d[i] - o->n_u];
      w = std::min(w, q);
    }
  }
  B->iluF.setup(o->A00, B->u_owner, n_parts);
  B->iluS.setup(o->S, B->p_owner, n_parts);
  return 0;
}

// reference :133-330 with every rank on its own cells (:166) and compress(add) (:292-294)
void nso_baseline_assemble(nso *o, double time) {
  NsoBaseline *B = o->baseline;
  const int dpc = o->dpc;
  const uint32_t nu_ = o->n_u;
  std::fill(o->A00.val.begin(), o->A00.val.end(), 0.0);
  std::fill(o->A01.val.begin(), o->A01.val.end(), 0.0);
  std::fill(o->A10.val.begin(), o->A10.val.end(), 0.0);
  std::fill(o->rhs.begin(), o->rhs.end(), 0.0);
  std::fill(o->lumped.begin(), o->lumped.end(), 0.0);
#pragma omp parallel num_threads(B->P)
  {
    CellFE fe;
    std::vector<double> cm((size_t)dpc * dpc), cr(dpc), cl(dpc);
#pragma omp for schedule(static, 1)
    for (int q = 0; q < B->P; ++q)
      for (int64_t c : B->cells_of[q]) {
        nso_cell_contribution(o, c, fe, cm.data(), cr.data(), cl.data());
        const uint32_t *dofs = &o->cell_dofs[(size_t)c * dpc];
        for (int i = 0; i < dpc; ++i)
          for (int j = 0; j < dpc; ++j) {
            const double v = cm[(size_t)i * dpc + j];
            if (v == 0.0) continue;
            const bool pi = dofs[i] >= nu_, pj = dofs[j] >= nu_;
            CsrMat &A = !pi ? (!pj ? o->A00 : o->A01) : o->A10;
            double &dst = A.val[A.find(pi ? dofs[i] - nu_ : dofs[i], pj ? dofs[j] - nu_ : dofs[j])];
#pragma omp atomic
            dst += v;
          }
        for (int i = 0; i < dpc; ++i) {
#pragma omp atomic
          o->rhs[dofs[i]] += cr[i];
#pragma omp atomic
          o->lumped[dofs[i]] += cl[i];
        }
      }
  }
  for (auto &x : o->lumped) x = o->deltat / x;
  nso_apply_boundary(o, time);
}

// reference :344-397 + PreconditionASIMPLE :934-995 on P ranks
int nso_baseline_solve_time_step(nso *o, int *iters, double *t_prec, double *t_solve) {
  NsoBaseline *B = o->baseline;
  const int P = B->P;
  g_threads = P;
  const uint32_t nu_ = o->n_u, np_ = o->n_p;
  const size_t N = (size_t)nu_ + np_;
  auto t0 = std::chrono::high_resolution_clock::now();
  const double tol = o->outer_rtol * pl2(o->rhs);  // :348
  Vec Di(nu_);
#pragma omp parallel for schedule(static) num_threads(P)
  for (int64_t i = 0; i < (int64_t)nu_; ++i) Di[i] = 1.0 / o->A00.val[o->A00.find(i, (uint32_t)i)];  // :948-953
  // :956 S = B diag(Di) Bt, every rank its own rows
#pragma omp parallel for schedule(dynamic, 64) num_threads(P)
  for (int64_t i = 0; i < (int64_t)np_; ++i) {
    for (int64_t k = o->S.rowptr[i]; k < o->S.rowptr[i + 1]; ++k) o->S.val[k] = 0.0;
    for (int64_t k = o->A10.rowptr[i]; k < o->A10.rowptr[i + 1]; ++k) {
      const uint32_t u = o->A10.colind[k];
      const double bu = o->A10.val[k] * Di[u];
      for (int64_t kk = o->A01.rowptr[u]; kk < o->A01.rowptr[u + 1]; ++kk)
        o->S.val[o->S.find(i, o->A01.colind[kk])] += bu * o->A01.val[kk];
    }
  }
  B->iluF.factor(o->A00, P);  // :958, rank-local
  B->iluS.factor(o->S, P);    // :959
  Vec vec0(nu_, 0.0), vec1(np_, 0.0);
  auto t1 = std::chrono::high_resolution_clock::now();

  ParGmres innerF, innerS, outer;
  innerF.max_it = innerS.max_it = 10000;
  bool inner_failed = false;
  long dbgF = 0, dbgS = 0;
  Vec src0(nu_), src1(np_), d0(nu_), d1(np_), tu(nu_);
  auto Pvmult = [&](Vec &dst, const Vec &src) {  // :966-995
    std::copy(src.begin(), src.begin() + nu_, src0.begin());
    std::copy(src.begin() + nu_, src.end(), src1.begin());
    std::copy(dst.begin() + nu_, dst.end(), d1.begin());  // stale initial guess
    innerF.tol = o->inner_rtol * pl2(src0);
    innerF.solve(nu_, [&](Vec &y, const Vec &x) { pvmult(o->A00, y.data(), x.data()); }, vec0, src0,
                 [&](Vec &y, const Vec &x) { B->iluF.vmult(y, x, P); });
    pvmult(o->A10, vec1.data(), vec0.data());
    for (uint32_t i = 0; i < np_; ++i) vec1[i] = -vec1[i] + src1[i];
    innerS.tol = o->inner_rtol * pl2(vec1);
    innerS.solve(np_, [&](Vec &y, const Vec &x) { pvmult(o->S, y.data(), x.data()); }, d1, vec1,
                 [&](Vec &y, const Vec &x) { B->iluS.vmult(y, x, P); });
    inner_failed = inner_failed || innerF.failed || innerS.failed;
    dbgF += innerF.last_step;
    dbgS += innerS.last_step;
    for (auto &x : d1) x *= -1.0 / o->alpha;
    pvmult(o->A01, d0.data(), d1.data());
#pragma omp parallel for schedule(static) num_threads(P)
    for (int64_t i = 0; i < (int64_t)nu_; ++i) d0[i] = -(d0[i] * Di[i]) + vec0[i];
    std::copy(d0.begin(), d0.end(), dst.begin());
    std::copy(d1.begin(), d1.end(), dst.begin() + nu_);
  };
  outer.n_tmp = o->n_tmp;
  outer.max_it = o->max_it;
  outer.tol = tol;
  outer.solve(N,
              [&](Vec &y, const Vec &x) {
                pvmult(o->A00, y.data(), x.data());
                pvmult(o->A01, tu.data(), x.data() + nu_);
#pragma omp parallel for schedule(static) num_threads(P)
                for (int64_t i = 0; i < (int64_t)nu_; ++i) y[i] += tu[i];
                pvmult(o->A10, y.data() + nu_, x.data());
              },
              o->solution_owned, o->rhs, Pvmult);
  auto t2 = std::chrono::high_resolution_clock::now();
  if (std::getenv("NSO_BASELINE_DEBUG"))
    std::fprintf(stderr, "baseline: outer %d, inner F its %ld, inner S its %ld\n", outer.last_step, dbgF, dbgS);
  if (iters) *iters = outer.last_step;
  if (t_prec) *t_prec = std::chrono::duration<double>(t1 - t0).count();
  if (t_solve) *t_solve = std::chrono::duration<double>(t2 - t1).count();
  o->solution = o->solution_owned;
  return (outer.failed || inner_failed) ? 1 : 0;
}

}  // extern "C"
