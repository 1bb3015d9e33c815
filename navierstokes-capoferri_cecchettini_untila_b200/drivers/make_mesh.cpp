// Writes a gmsh .msh (format 2.2) of one of the generated meshes, standing in
// for `gmsh mesh/*.geo` (gmsh is not available here).
//   make_mesh <2d-cylinder|3d-square|3d-cylinder|naca2412|channel2d|channel3d> <h> <out.msh>
#include <cstdlib>
#include <iostream>

#include "mesh.hpp"

int main(int argc, char **argv) {
  if (argc != 4) {
    std::cerr << "usage: make_mesh <name> <h> <out.msh>\n";
    return 2;
  }
  try {
    const nsb::Mesh m = nsb::gen_named(argv[1], std::atof(argv[2]));
    nsb::write_msh(m, argv[3]);
    std::cout << argv[3] << ": " << m.n_verts() << " vertices, " << m.n_cells() << " cells, " << m.n_bfaces()
              << " boundary facets\n";
  } catch (const std::exception &e) {
    std::cerr << "make_mesh: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
