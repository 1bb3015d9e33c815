// Drag / lift integrals over the faces with boundary id 4 (replaces reference
// src/NavierStokes.cpp:859-909).  The reference evaluates pressure and velocity
// gradients at the CELL quadrature points and indexes them with the FACE
// quadrature counter; that is reproduced here: point q of the
// cell rule, q < n_q_face, is paired with face weight q, the face normal and
// the face measure.  Only boundary-4 faces are visited (the reference loops
// over every cell but uses nothing else).
#pragma once
#include "common.cuh"

namespace nsb {

template <int DIM>
__global__ void __launch_bounds__(128) forces_kernel(int64_t n_faces, const uint32_t *__restrict__ face_cell,
                                                     const double *__restrict__ normal,
                                                     const double *__restrict__ measure,
                                                     const double *__restrict__ xyz,
                                                     const uint32_t *__restrict__ cell_verts,
                                                     const uint32_t *__restrict__ cell_nodes,
                                                     const uint32_t *__restrict__ cell_pverts,
                                                     const double *__restrict__ sol, int64_t p_offset /* start of the pressure part */,
                                                     const FeTables *__restrict__ fe, double nu,
                                                     double *__restrict__ out /* drag, lift */) {
  constexpr int NV = DIM + 1, NN = DIM == 2 ? 6 : 10;
  double drag = 0, lift = 0;
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_faces;
       f += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = face_cell[f];
    double X[NV][DIM];
    for (int a = 0; a < NV; ++a)
      for (int r = 0; r < DIM; ++r) X[a][r] = xyz[(size_t)cell_verts[cell * NV + a] * DIM + r];
    double J[DIM][DIM], Ji[DIM][DIM];
    for (int a = 0; a < DIM; ++a)
      for (int r = 0; r < DIM; ++r) J[r][a] = X[a + 1][r] - X[0][r];
    if constexpr (DIM == 2) {
      const double id = 1.0 / (J[0][0] * J[1][1] - J[0][1] * J[1][0]);
      Ji[0][0] = J[1][1] * id;
      Ji[0][1] = -J[0][1] * id;
      Ji[1][0] = -J[1][0] * id;
      Ji[1][1] = J[0][0] * id;
    } else {
      const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      const double id = 1.0 / (J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02);
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
    double n[3] = {0, 0, 0};
    for (int r = 0; r < DIM; ++r) n[r] = normal[f * DIM + r];
    const double tangent[3] = {n[1], -n[0], 0.0};  // reference :886-890
    const double meas = measure[f];
    for (int q = 0; q < fe->nqf; ++q) {
      double p = 0, ngt = 0;
      for (int k = 0; k < NV; ++k) p += sol[p_offset + cell_pverts[cell * NV + k]] * fe->psi[q][k];
      for (int a = 0; a < NN; ++a) {
        double g[DIM];
        for (int c = 0; c < DIM; ++c) {
          g[c] = 0;
          for (int d = 0; d < DIM; ++d) g[c] += Ji[d][c] * fe->dphi[q][a][d];
        }
        double gt = 0, un = 0;  // (grad phi_a . tangent), (U_a . normal)
        for (int c = 0; c < DIM; ++c) {
          gt += g[c] * tangent[c];
          un += sol[(size_t)DIM * cell_nodes[cell * NN + a] + c] * n[c];
        }
        ngt += un * gt;  // n_i (d u_i / d x_j) t_j
      }
      const double JxW = fe->wface[q] * meas;
      drag += nu * ngt * n[1] * JxW - p * n[0] * JxW;   // reference :892-896
      lift += -nu * ngt * n[0] * JxW - p * n[1] * JxW;  // reference :898-902
    }
  }
  __shared__ double s_d[4], s_l[4];
  for (int o = 16; o > 0; o >>= 1) {
    drag += __shfl_xor_sync(0xffffffffu, drag, o);
    lift += __shfl_xor_sync(0xffffffffu, lift, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_d[threadIdx.x >> 5] = drag;
    s_l[threadIdx.x >> 5] = lift;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0, l = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      d += s_d[i];
      l += s_l[i];
    }
    atomicAdd(out + 0, d);
    atomicAdd(out + 1, l);
  }
}

}  // namespace nsb
