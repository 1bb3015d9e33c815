// Simplex mesh container, gmsh .msh I/O and deterministic mesh generators.
//
// Replaces what the reference obtains from deal.II's GridIn::read_msh
// (reference src/NavierStokes.cpp:11-17) and from gmsh runs over mesh/*.geo
// (the .msh files are not in the reference tree, .gitignore:39).  Semantics
// follow SURVEY.md Appendix A.1: vertices in file order, cells in file order,
// boundary ids from the lower-dimensional elements' physical tags, untagged
// boundary facets keep id 0.
#pragma once
#include <cstdint>
#include <array>
#include <string>
#include <vector>

namespace nsb {

struct Mesh {
  int dim = 0;                      // 2 (triangles) or 3 (tetrahedra)
  std::vector<double> xyz;          // n_verts * dim
  std::vector<uint32_t> cells;      // n_cells * (dim+1), positively oriented
  std::vector<uint32_t> bfaces;     // n_bfaces * dim (vertex ids of tagged boundary facets)
  std::vector<int32_t> bids;        // n_bfaces physical tags
  size_t n_verts() const { return dim ? xyz.size() / dim : 0; }
  size_t n_cells() const { return dim ? cells.size() / (dim + 1) : 0; }
  size_t n_bfaces() const { return dim ? bfaces.size() / dim : 0; }
};

// gmsh ASCII reader (format 2.2 and 4.1; element types 1, 2, 4, 15).  `dim`
// selects which element type is the cell (2: triangles, 3: tetrahedra).
// Throws std::runtime_error on a missing/ill-formed file (the reference's
// read_msh throws too, SURVEY.md §8b "Error convention").
Mesh read_msh(const std::string &path, int dim);
// gmsh ASCII 2.2 writer: physical tag 10 on cells (mesh/domain2D.geo:44),
// boundary facets carry their id as the physical tag.
void write_msh(const Mesh &m, const std::string &path);

// Makes every cell positively oriented (swaps the last two vertices otherwise)
// and returns the number of cells that were flipped.
size_t orient_cells(Mesh &m);

// ---- generators (deterministic; no RNG) ---------------------------------
// 2D channel [0,Lx]x[0,Ly] with a circular hole: an O-grid of straight rays
// between the circle and a square box of side Ly around it, plus structured
// up/downstream blocks.  Boundary ids as in mesh/domain2D.geo:39-43:
// 0 bottom, 1 outlet, 2 top, 3 inlet, 4 obstacle.
Mesh gen_channel2d_circle(double Lx, double Ly, double cx, double cy, double r,
                          double h);
// 2D channel with a square hole [ox,ox+s]x[oy,oy+s] on a Cartesian grid
// (cross-section of mesh/domain3D.geo:2-12).  Same boundary ids.
Mesh gen_channel2d_square(double Lx, double Ly, double ox, double oy, double s,
                          double h);
// Plain channel without obstacle (Poiseuille KAT).
Mesh gen_channel2d_plain(double Lx, double Ly, int nx, int ny);
// NACA 4-digit airfoil (chord 1, nose at (cx-0.5, cy)) in a far-field box,
// rotated by `aoa_deg` about (cx,cy) like mesh/test.py:25-41.  O-grid.
Mesh gen_naca2d(double Lx, double Ly, double cx, double cy, int naca4,
                double aoa_deg, double chord, int n_around, int n_radial);
// The reference's airfoil pre-processing (mesh/test.py:25-41, 155-168): contour points with the chord on
// [0,1] (read_airfoil_dat: the mesh/naca.dat layout, or naca4_contour) are shifted to mid-chord, scaled to
// `chord`, turned clockwise by the angle of attack and placed at (cx, cy); gen_airfoil2d meshes the box around
// that closed, counter-clockwise contour (which must be star-shaped with respect to (cx, cy)).
std::vector<std::array<double, 2>> read_airfoil_dat(const std::string &path, std::string *name = nullptr);
std::vector<std::array<double, 2>> naca4_contour(int naca4, int n_around);
std::vector<std::array<double, 2>> place_airfoil(const std::vector<std::array<double, 2>> &unit, double chord,
                                                 double aoa_deg, double cx, double cy);
Mesh gen_airfoil2d(double Lx, double Ly, double cx, double cy, const std::vector<std::array<double, 2>> &contour,
                   double chord, int n_radial);
// Extrudes a triangle mesh in z into nz layers of prisms, each split into 3
// tetrahedra with the smallest-vertex-index diagonal rule (conforming).
// 3D boundary ids as in mesh/domain3D.geo:104-108: z-planes 0, outlet 1,
// y-walls (2D ids 0 and 2) 2, inlet 3, obstacle 4.
Mesh extrude_to_tets(const Mesh &m2, double Lz, int nz);

// The five BASELINE.json configs by name: "2d-cylinder", "3d-square",
// "3d-cylinder", "naca2412", "channel2d", "channel3d".  `h` is the target
// edge length.
Mesh gen_named(const std::string &name, double h);

}  // namespace nsb
