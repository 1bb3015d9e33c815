// Writes a gmsh .msh (format 2.2) of one of the generated meshes, standing in
// for `gmsh mesh/*.geo` (gmsh is not available here).
//   make_mesh <2d-cylinder|3d-square|3d-cylinder|naca2412|channel2d|channel3d> <h> <out.msh>
//   make_mesh airfoil <contour.dat | nacaDDDD> <chord> <angle of attack, degrees> <h> <out.msh>
// The second form is the reference's `./test.py naca.dat 0.4 <angle>; gmsh NACA_2408.geo -2 -o domain2D.msh`
// (tests/2D/test_naca/run_test.sh:7-9, mesh/test.py:25-41, 155-168) in test.py's 2.2 x 1.0 box, centre (0.4, 0.5).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <iostream>

#include "mesh.hpp"

int main(int argc, char **argv) {
  const bool airfoil = argc == 7 && std::string(argv[1]) == "airfoil";
  if (argc != 4 && !airfoil) {
    std::cerr << "usage: make_mesh <name> <h> <out.msh>\n"
                 "       make_mesh airfoil <contour.dat|nacaDDDD> <chord> <aoa_deg> <h> <out.msh>\n";
    return 2;
  }
  try {
    nsb::Mesh m;
    const char *out = argv[3];
    if (airfoil) {
      const std::string src = argv[2];
      const double chord = std::atof(argv[3]), aoa = std::atof(argv[4]), h = std::atof(argv[5]);
      out = argv[6];
      const int n_around = std::max(32, (int)std::lround(2.1 * chord / h));
      const auto unit = src.rfind("naca", 0) == 0 && src.find('.') == std::string::npos
                            ? nsb::naca4_contour(std::atoi(src.c_str() + 4), n_around)
                            : nsb::read_airfoil_dat(src);
      m = nsb::gen_airfoil2d(2.2, 1.0, 0.4, 0.5, nsb::place_airfoil(unit, chord, aoa, 0.4, 0.5), chord, 48);
    } else
      m = nsb::gen_named(argv[1], std::atof(argv[2]));
    nsb::write_msh(m, out);
    std::cout << out << ": " << m.n_verts() << " vertices, " << m.n_cells() << " cells, " << m.n_bfaces()
              << " boundary facets\n";
  } catch (const std::exception &e) {
    std::cerr << "make_mesh: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
