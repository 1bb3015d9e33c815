// Shared body of the driver mains: parameters per case follow the reference's
// test drivers.  Each driver defines NS_INPUT, includes
// NavierStokes.hpp and supplies InletVelocity in the reference's way.
#pragma once
#include <cstdlib>
#include <string>

inline int run_case(int argc, char **argv, const std::string &default_mesh, double deltat, double T,
                    unsigned int out_step, int Re /* <= 0: keep nu = 1e-3 */) {
  Utilities::MPI::MPI_InitFinalize mpi_init(argc, argv);
  const std::string mesh_file_name = argc > 1 ? argv[1] : default_mesh;
  if (argc > 2) T = std::atof(argv[2]);  // optional shorter end time for smoke runs
  NavierStokes problem(mesh_file_name, 2, 1, deltat, T, out_step);
  if (Re > 0) problem.set_re_number(Re);
  problem.setup();
  problem.compute_ordered_dofs_indices();
  problem.solve();
  return 0;
}
