// Quadrature-point tables of the Taylor-Hood P2/P1 pair on the reference
// simplex, and the geometry-independent contractions the assembly kernel uses.
//
// Stands in for what deal.II's FEValues evaluates per cell in the reference
// (src/NavierStokes.cpp:141-146, 169): because MappingFE(FE_SimplexP(1)) is
// affine, shape values at the quadrature points are the same
// for every cell and physical gradients are J^{-T} times the reference ones.
#pragma once
#include <cmath>
#include <cstring>

namespace nsb {

constexpr int kMaxQ = 14, kMaxNN = 10, kMaxNV = 4;

struct FeTables {
  int dim, nq, nn, nv, nqf, pad_[3];
  double w[kMaxQ];                     // cell weights (sum = reference volume)
  double phi[kMaxQ][kMaxNN];           // P2 values
  double dphi[kMaxQ][kMaxNN][3];       // P2 reference gradients
  double psi[kMaxQ][kMaxNV];           // P1 values
  double mhat[kMaxNN][kMaxNN];         // sum_q w phi_a phi_b
  double dhat[kMaxNN][kMaxNV][3];      // sum_q w d_d(phi_a) psi_k
  double wface[7];                     // face weights, normalised to sum 1
  double labs[kMaxNN];                 // sum_q w |phi_a| sum_b |phi_b|  (lumped mass of reference :232-236)
  // geometry-independent contractions of the stiffness and convection terms (pair p = a*nn + b):
  //   khat[d*dim+e][p] = sum_q w d_d(phi_a) d_e(phi_b)
  //   chat[n*dim+d][p] = sum_q w phi_a phi_n d_d(phi_b)
  double khat[9][kMaxNN * kMaxNN];
  double chat[kMaxNN * 3][kMaxNN * kMaxNN];
};

namespace fe_detail {
// reference-cell line -> vertex tables (deal.II ReferenceCell order)
static const int tri_lines[3][2] = {{0, 1}, {1, 2}, {2, 0}};
static const int tet_lines[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};

inline void barycentric(int dim, const double *x, double *lam) {
  lam[0] = 1.0;
  for (int d = 0; d < dim; ++d) {
    lam[0] -= x[d];
    lam[d + 1] = x[d];
  }
}
}  // namespace fe_detail

// rule: 0 = deal.II 9.3.x tables, 1 = deal.II >= 9.4 (Witherden-Vincent),
// see include/nsb.h and SURVEY.md A.4 / H2.
inline bool fill_fe_tables(int dim, int rule, FeTables &T) {
  using namespace fe_detail;
  std::memset(&T, 0, sizeof(T));
  T.dim = dim;
  T.nv = dim + 1;
  T.nn = dim == 2 ? 6 : 10;
  double pts[kMaxQ][3] = {{0}};
  if (dim == 2) {
    T.nq = 7;
    const double r15 = std::sqrt(15.0);
    const double lo = (6.0 - r15) / 21.0, hi = 1.0 - 2.0 * lo;     // 0.1012.., 0.7974..
    const double mid = (6.0 + r15) / 21.0, sm = 1.0 - 2.0 * mid;   // 0.4701.., 0.0597..
    if (rule == 0) {
      const double P[7][2] = {{0.3333333333330, 0.3333333333330}, {0.7974269853530, 0.1012865073230},
                              {0.1012865073230, 0.7974269853530}, {0.1012865073230, 0.1012865073230},
                              {0.0597158717898, 0.4701420641050}, {0.4701420641050, 0.0597158717898},
                              {0.4701420641050, 0.4701420641050}};
      const double W[7] = {0.225, 0.125939180545, 0.125939180545, 0.125939180545,
                           0.132394152789, 0.132394152789, 0.132394152789};
      for (int q = 0; q < 7; ++q) {
        pts[q][0] = P[q][0];
        pts[q][1] = P[q][1];
        T.w[q] = 0.5 * W[q];
      }
    } else {
      const double P[7][2] = {{1.0 / 3.0, 1.0 / 3.0}, {lo, lo}, {lo, hi}, {hi, lo}, {sm, mid}, {mid, sm}, {mid, mid}};
      const double wa = (155.0 - r15) / 1200.0, wb = (155.0 + r15) / 1200.0;
      const double W[7] = {9.0 / 40.0, wa, wa, wa, wb, wb, wb};
      for (int q = 0; q < 7; ++q) {
        pts[q][0] = P[q][0];
        pts[q][1] = P[q][1];
        T.w[q] = 0.5 * W[q];
      }
    }
    T.nqf = 3;
    T.wface[0] = 5.0 / 18.0;
    T.wface[1] = 8.0 / 18.0;
    T.wface[2] = 5.0 / 18.0;
  } else if (dim == 3) {
    if (rule == 0) {
      T.nq = 10;
      const double a = 0.5684305841968444, b = 0.1438564719343852;
      const double P[10][3] = {{a, b, b},   {b, b, b},   {b, b, a},   {b, a, b},   {0, .5, .5},
                               {.5, 0, .5}, {.5, .5, 0}, {.5, 0, 0}, {0, .5, 0}, {0, 0, .5}};
      for (int q = 0; q < 10; ++q) {
        for (int d = 0; d < 3; ++d) pts[q][d] = P[q][d];
        T.w[q] = (q < 4 ? 0.2177650698804054 : 0.0214899534130631) / 6.0;
      }
    } else {
      T.nq = 14;
      // orbit parameters: two S31 orbits and one S22 orbit, each listed in the
      // lexicographic permutation order of its sorted barycentric 4-tuple
      const double s1 = 3.1088591926330061e-01, t1 = 1.0 - 3.0 * s1;
      const double s2 = 9.2735250310891248e-02, t2 = 1.0 - 3.0 * s2;
      const double s3 = 4.5503704125649642e-02, t3 = 0.5 * (1.0 - 2.0 * s3);
      const double P[14][3] = {{t1, s1, s1}, {s1, t1, s1}, {s1, s1, t1}, {s1, s1, s1}, {s2, s2, s2},
                               {s2, s2, t2}, {s2, t2, s2}, {t2, s2, s2}, {s3, s3, t3}, {s3, t3, s3},
                               {s3, t3, t3}, {t3, s3, s3}, {t3, s3, t3}, {t3, t3, s3}};
      for (int q = 0; q < 14; ++q) {
        for (int d = 0; d < 3; ++d) pts[q][d] = P[q][d];
        T.w[q] = (q < 4 ? 1.1268792571801590e-01 : q < 8 ? 7.3493043116361956e-02 : 4.2546020777081472e-02) / 6.0;
      }
    }
    // face rule = the 2D 7-point rule; only its weights enter (face JxW)
    FeTables F2;
    fill_fe_tables(2, rule, F2);
    T.nqf = 7;
    for (int q = 0; q < 7; ++q) T.wface[q] = 2.0 * F2.w[q];
  } else
    return false;

  const int nl = dim == 2 ? 3 : 6;
  for (int q = 0; q < T.nq; ++q) {
    double lam[4];
    barycentric(dim, pts[q], lam);
    // d(lam_a)/d(xhat_d): -1 for a = 0, delta(a-1, d) otherwise
    auto dlam = [&](int a, int d) { return a == 0 ? -1.0 : (a - 1 == d ? 1.0 : 0.0); };
    for (int a = 0; a < T.nv; ++a) {
      T.phi[q][a] = lam[a] * (2.0 * lam[a] - 1.0);
      T.psi[q][a] = lam[a];
      for (int d = 0; d < dim; ++d) T.dphi[q][a][d] = (4.0 * lam[a] - 1.0) * dlam(a, d);
    }
    for (int l = 0; l < nl; ++l) {
      const int i = dim == 2 ? tri_lines[l][0] : tet_lines[l][0];
      const int j = dim == 2 ? tri_lines[l][1] : tet_lines[l][1];
      T.phi[q][T.nv + l] = 4.0 * lam[i] * lam[j];
      for (int d = 0; d < dim; ++d) T.dphi[q][T.nv + l][d] = 4.0 * (lam[i] * dlam(j, d) + lam[j] * dlam(i, d));
    }
  }
  for (int a = 0; a < T.nn; ++a) {
    for (int b = 0; b < T.nn; ++b) {
      double s = 0;
      for (int q = 0; q < T.nq; ++q) s += T.w[q] * T.phi[q][a] * T.phi[q][b];
      T.mhat[a][b] = s;
    }
    for (int b = 0; b < T.nn; ++b) {
      const int p = a * T.nn + b;
      for (int d = 0; d < dim; ++d)
        for (int e = 0; e < dim; ++e) {
          double s = 0;
          for (int q = 0; q < T.nq; ++q) s += T.w[q] * T.dphi[q][a][d] * T.dphi[q][b][e];
          T.khat[d * dim + e][p] = s;
        }
      for (int n = 0; n < T.nn; ++n)
        for (int d = 0; d < dim; ++d) {
          double s = 0;
          for (int q = 0; q < T.nq; ++q) s += T.w[q] * T.phi[q][a] * T.phi[q][n] * T.dphi[q][b][d];
          T.chat[n * dim + d][p] = s;
        }
    }
    {
      double s = 0;  // the absolute value is taken per (q, j) term, as in the reference
      for (int q = 0; q < T.nq; ++q)
        for (int b = 0; b < T.nn; ++b) s += std::fabs(T.w[q] * T.phi[q][a] * T.phi[q][b]);
      T.labs[a] = s;
    }
    for (int k = 0; k < T.nv; ++k)
      for (int d = 0; d < dim; ++d) {
        double s = 0;
        for (int q = 0; q < T.nq; ++q) s += T.w[q] * T.dphi[q][a][d] * T.psi[q][k];
        T.dhat[a][k][d] = s;
      }
  }
  return true;
}

}  // namespace nsb
