// One discretised problem: mesh + Taylor-Hood space + boundary data.  This is
// the C++ object behind `nsh_problem` (include/nsb_host.h) and the state the
// NavierStokes facade keeps between setup() and the time loop.
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "fespace.hpp"
#include "mesh.hpp"

namespace nsb {

// The drivers' InletVelocity (reference tests/*/src/*.cpp:18-42): a profile in
// component 0, optionally modulated by sin(pi t / 8) (test_03 drivers).
struct Inlet {
  int kind = 0;  // 0 parabolic, 1 uniform
  double U_m = 0.3, H = 0.41;
  int time_sin = 0;
  double profile(int dim, const double *p, int comp) const {
    if (comp != 0) return 0.0;
    if (kind == 1) return U_m;
    if (dim == 2) return 4 * U_m * p[1] * (H - p[1]) / (H * H);
    return 16 * U_m * p[1] * p[2] * (H - p[1]) * (H - p[2]) / (H * H * H * H);
  }
  double time_factor(double t) const { return time_sin ? std::sin(M_PI * t / 8.0) : 1.0; }
  double mean_vel(int dim, double t) const {  // get_mean_vel()
    const double v = kind == 1 ? U_m : (dim == 2 ? 2.0 * U_m / 3.0 : 4.0 * U_m / 9.0);
    return v * time_factor(t);
  }
};

struct Problem {
  Mesh mesh;
  DofMap dofs;
  Patterns pat;
  Inlet inlet;
  std::vector<BoundaryFace> bfaces;
  DirichletSet bc;
  ForceFaces ff;
  std::vector<int32_t> part_cell;  // cell -> part (after partition())
  bool has_space = false, has_boundary = false;

  void build_space(bool expand_a00 = true) {
    dofs = build_dofmap(mesh);
    pat = build_patterns(mesh, dofs, expand_a00);
    has_space = true;
  }
  void build_boundary() {
    bfaces = boundary_faces(mesh);
    const int dim = mesh.dim;
    bc = dirichlet_dofs(mesh, dofs, bfaces, [&](const double *p, int c) { return inlet.profile(dim, p, c); });
    ff = force_faces(mesh, bfaces, 4);
    has_boundary = true;
  }
  void partition(int n_parts);
};

}  // namespace nsb
