// Restart / post-processing driver (tests/test_facade.py): the paths of the reference that its own drivers
// leave to the user -- solve(time_step != 0) with import_data (reference src/NavierStokes.cpp:457-463, 787-805)
// and post_process (:808-828, src/postprocess.cpp:1-19) -- on the d3_test_01 parameters, output step 1.
//   d3_restart <mesh> <T> full            solve() from t = 0, a checkpoint per step
//   d3_restart <mesh> <T> restart <k>     solve(k): continue from ../cache/state-ns-<k>.dat
//   d3_restart <mesh> <T> post <k0> <k1>  post_process(k0, k1, 1)
#define NS_INPUT
#include "NavierStokes.hpp"

#include <cstdlib>
#include <string>

static constexpr double U_m = 0.45;
static constexpr double H = 0.41;

double NavierStokes::InletVelocity::value(const Point<dim> &p, const unsigned int component) const {
  (void)p;
  return component == 0 ? (16 * U_m * p[1] * p[2] * (H - p[1]) * (H - p[2]) / (H * H * H * H)) : 0.0;
}
void NavierStokes::InletVelocity::vector_value(const Point<dim> &p, Vector<double> &values) const {
  for (unsigned int i = 0; i < dim + 1; ++i) values[i] = value(p, i);
}
double NavierStokes::InletVelocity::get_mean_vel() { return 4.0 * U_m / 9.0; }

int main(int argc, char **argv) {
  Utilities::MPI::MPI_InitFinalize mpi_init(argc, argv);
  if (argc < 4) return 2;
  const std::string mode = argv[3];
  NavierStokes problem(argv[1], 2, 1, 0.01, std::atof(argv[2]), 1);
  problem.set_re_number(20);
  problem.setup();
  problem.compute_ordered_dofs_indices();
  if (mode == "full")
    problem.solve();
  else if (mode == "restart" && argc > 4)
    problem.solve((unsigned int)std::atoi(argv[4]));
  else if (mode == "post" && argc > 5)
    problem.post_process((unsigned int)std::atoi(argv[4]), (unsigned int)std::atoi(argv[5]), 1);
  else
    return 2;
  return 0;
}
