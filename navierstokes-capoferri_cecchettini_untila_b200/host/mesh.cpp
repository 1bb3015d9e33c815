// gmsh .msh I/O and deterministic generators.  See mesh.hpp.
#include "mesh.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace nsb {

namespace {

double cell_det(const Mesh &m, size_t c) {
  const int d = m.dim;
  const uint32_t *v = &m.cells[c * (d + 1)];
  const double *p0 = &m.xyz[(size_t)v[0] * d];
  double J[3][3];
  for (int a = 0; a < d; ++a) {
    const double *pa = &m.xyz[(size_t)v[a + 1] * d];
    for (int r = 0; r < d; ++r) J[r][a] = pa[r] - p0[r];
  }
  if (d == 2) return J[0][0] * J[1][1] - J[0][1] * J[1][0];
  return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
         J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
         J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
}

// Removes vertices no cell refers to (deal.II's read_msh does the same,
// SURVEY.md A.1) keeping the relative order of the survivors.
void drop_unused_vertices(Mesh &m) {
  const size_t nv = m.n_verts();
  std::vector<uint8_t> used(nv, 0);
  for (uint32_t v : m.cells) used[v] = 1;
  std::vector<uint32_t> remap(nv, UINT32_MAX);
  size_t k = 0;
  for (size_t v = 0; v < nv; ++v)
    if (used[v]) {
      remap[v] = (uint32_t)k;
      if (k != v)
        for (int r = 0; r < m.dim; ++r) m.xyz[k * m.dim + r] = m.xyz[v * m.dim + r];
      ++k;
    }
  if (k == nv) return;
  m.xyz.resize(k * m.dim);
  for (auto &v : m.cells) v = remap[v];
  // boundary facets touching a dropped vertex cannot belong to a cell
  std::vector<uint32_t> bf;
  std::vector<int32_t> bi;
  for (size_t f = 0; f < m.n_bfaces(); ++f) {
    bool ok = true;
    for (int r = 0; r < m.dim; ++r) ok = ok && remap[m.bfaces[f * m.dim + r]] != UINT32_MAX;
    if (!ok) continue;
    for (int r = 0; r < m.dim; ++r) bf.push_back(remap[m.bfaces[f * m.dim + r]]);
    bi.push_back(m.bids[f]);
  }
  m.bfaces.swap(bf);
  m.bids.swap(bi);
}

// Quad (a,b,c,d) counter-clockwise -> two CCW triangles, diagonal from the
// smallest vertex index.
void push_quad(std::vector<uint32_t> &tris, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  uint32_t q[4] = {a, b, c, d};
  int s = 0;
  for (int i = 1; i < 4; ++i)
    if (q[i] < q[s]) s = i;
  uint32_t r[4] = {q[s], q[(s + 1) & 3], q[(s + 2) & 3], q[(s + 3) & 3]};
  tris.insert(tris.end(), {r[0], r[1], r[2], r[0], r[2], r[3]});
}

// Normalised graded abscissae s_0=0..s_n=1 with spacings growing
// geometrically from ~d0 to ~d1 over a total length len.
std::vector<double> graded(double len, double d0, double d1, int *n_out) {
  d0 = std::min(d0, len);
  d1 = std::max(d1, d0);
  int n = std::max(2, (int)std::lround(len / (0.5 * (d0 + d1))));
  double g = std::pow(d1 / d0, 1.0 / (n - 1));
  std::vector<double> s(n + 1, 0.0);
  double w = 1.0, acc = 0.0;
  for (int i = 0; i < n; ++i) {
    acc += w;
    s[i + 1] = acc;
    w *= g;
  }
  for (auto &x : s) x /= acc;
  s[n] = 1.0;
  *n_out = n;
  return s;
}

}  // namespace

size_t orient_cells(Mesh &m) {
  size_t flipped = 0;
  const int d = m.dim;
  for (size_t c = 0; c < m.n_cells(); ++c)
    if (cell_det(m, c) < 0) {
      std::swap(m.cells[c * (d + 1) + d - 1], m.cells[c * (d + 1) + d]);
      ++flipped;
    }
  return flipped;
}

// --------------------------------------------------------------------------
// gmsh I/O
// --------------------------------------------------------------------------
void write_msh(const Mesh &m, const std::string &path) {
  FILE *f = std::fopen(path.c_str(), "w");
  if (!f) throw std::runtime_error("write_msh: cannot open " + path);
  std::fprintf(f, "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%zu\n", m.n_verts());
  for (size_t v = 0; v < m.n_verts(); ++v) {
    const double *p = &m.xyz[v * m.dim];
    std::fprintf(f, "%zu %.17g %.17g %.17g\n", v + 1, p[0], p[1], m.dim == 3 ? p[2] : 0.0);
  }
  std::fprintf(f, "$EndNodes\n$Elements\n%zu\n", m.n_bfaces() + m.n_cells());
  size_t id = 1;
  const int ftype = m.dim == 2 ? 1 : 2, ctype = m.dim == 2 ? 2 : 4;
  for (size_t b = 0; b < m.n_bfaces(); ++b, ++id) {
    std::fprintf(f, "%zu %d 2 %d %d", id, ftype, m.bids[b], m.bids[b]);
    for (int r = 0; r < m.dim; ++r) std::fprintf(f, " %u", m.bfaces[b * m.dim + r] + 1);
    std::fputc('\n', f);
  }
  for (size_t c = 0; c < m.n_cells(); ++c, ++id) {
    std::fprintf(f, "%zu %d 2 10 1", id, ctype);
    for (int r = 0; r <= m.dim; ++r) std::fprintf(f, " %u", m.cells[c * (m.dim + 1) + r] + 1);
    std::fputc('\n', f);
  }
  std::fprintf(f, "$EndElements\n");
  std::fclose(f);
}

Mesh read_msh(const std::string &path, int dim) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("read_msh: cannot open " + path);
  Mesh m;
  m.dim = dim;
  const int ctype = dim == 2 ? 2 : 4, ftype = dim == 2 ? 1 : 2;
  static const int nodes_of_type[16] = {0, 2, 3, 4, 4, 8, 6, 5, 3, 6, 9, 10, 27, 18, 14, 1};
  double version = 0;
  std::unordered_map<long long, uint32_t> tag2idx;  // node tag -> vertex index
  // 4.1: (entity dim, entity tag) -> physical tag
  std::map<std::pair<int, long long>, int> ent_phys;
  std::string line;
  auto expect_end = [&](const char *what) {
    while (std::getline(in, line))
      if (line.rfind(what, 0) == 0) return;
    throw std::runtime_error(std::string("read_msh: missing ") + what);
  };
  auto add_element = [&](int type, int phys, const std::vector<long long> &tags) {
    if (type != ctype && type != ftype) return;
    std::vector<uint32_t> &dst = type == ctype ? m.cells : m.bfaces;
    for (long long t : tags) {
      auto it = tag2idx.find(t);
      if (it == tag2idx.end()) throw std::runtime_error("read_msh: element refers to unknown node");
      dst.push_back(it->second);
    }
    if (type == ftype) m.bids.push_back(phys);
  };
  while (std::getline(in, line)) {
    if (line.rfind("$MeshFormat", 0) == 0) {
      int ft, ds;
      in >> version >> ft >> ds;
      if (ft != 0) throw std::runtime_error("read_msh: binary .msh not supported");
      expect_end("$EndMeshFormat");
    } else if (line.rfind("$Entities", 0) == 0 && version >= 4.0) {
      size_t n[4];
      in >> n[0] >> n[1] >> n[2] >> n[3];
      for (int ed = 0; ed < 4; ++ed)
        for (size_t e = 0; e < n[ed]; ++e) {
          long long tag;
          in >> tag;
          double tmp;
          for (int k = 0; k < (ed == 0 ? 3 : 6); ++k) in >> tmp;
          size_t np;
          in >> np;
          for (size_t k = 0; k < np; ++k) {
            int p;
            in >> p;
            if (k == 0) ent_phys[{ed, tag}] = p;
          }
          if (ed > 0) {
            size_t nb;
            in >> nb;
            long long t;
            for (size_t k = 0; k < nb; ++k) in >> t;
          }
        }
      expect_end("$EndEntities");
    } else if (line.rfind("$Nodes", 0) == 0) {
      if (version >= 4.0) {
        size_t nblocks, nn, mn, mx;
        in >> nblocks >> nn >> mn >> mx;
        m.xyz.reserve(nn * dim);
        for (size_t b = 0; b < nblocks; ++b) {
          int ed, par;
          long long et;
          size_t cnt;
          in >> ed >> et >> par >> cnt;
          std::vector<long long> tags(cnt);
          for (auto &t : tags) in >> t;
          for (size_t k = 0; k < cnt; ++k) {
            double p[3];
            in >> p[0] >> p[1] >> p[2];
            tag2idx[tags[k]] = (uint32_t)(m.xyz.size() / dim);
            for (int r = 0; r < dim; ++r) m.xyz.push_back(p[r]);
          }
        }
      } else {
        size_t nn;
        in >> nn;
        m.xyz.reserve(nn * dim);
        for (size_t k = 0; k < nn; ++k) {
          long long t;
          double p[3];
          in >> t >> p[0] >> p[1] >> p[2];
          tag2idx[t] = (uint32_t)k;
          for (int r = 0; r < dim; ++r) m.xyz.push_back(p[r]);
        }
      }
      if (!in) throw std::runtime_error("read_msh: ill-formed $Nodes");
      expect_end("$EndNodes");
    } else if (line.rfind("$Elements", 0) == 0) {
      if (version >= 4.0) {
        size_t nblocks, ne, mn, mx;
        in >> nblocks >> ne >> mn >> mx;
        for (size_t b = 0; b < nblocks; ++b) {
          int ed, type;
          long long et;
          size_t cnt;
          in >> ed >> et >> type >> cnt;
          auto it = ent_phys.find({ed, et});
          const int phys = it == ent_phys.end() ? 0 : it->second;
          const int nn = (type >= 1 && type <= 15) ? nodes_of_type[type] : 0;
          if (!nn) throw std::runtime_error("read_msh: unsupported element type");
          std::vector<long long> tags(nn);
          for (size_t k = 0; k < cnt; ++k) {
            long long id;
            in >> id;
            for (auto &t : tags) in >> t;
            add_element(type, phys, tags);
          }
        }
      } else {
        size_t ne;
        in >> ne;
        for (size_t k = 0; k < ne; ++k) {
          long long id;
          int type, ntags;
          in >> id >> type >> ntags;
          int phys = 0;
          for (int t = 0; t < ntags; ++t) {
            int v;
            in >> v;
            if (t == 0) phys = v;
          }
          const int nn = (type >= 1 && type <= 15) ? nodes_of_type[type] : 0;
          if (!nn) throw std::runtime_error("read_msh: unsupported element type");
          std::vector<long long> tags(nn);
          for (auto &t : tags) in >> t;
          add_element(type, phys, tags);
        }
      }
      if (!in) throw std::runtime_error("read_msh: ill-formed $Elements");
      expect_end("$EndElements");
    }
  }
  if (m.cells.empty()) throw std::runtime_error("read_msh: no cells of the requested dimension in " + path);
  drop_unused_vertices(m);
  orient_cells(m);
  return m;
}

// --------------------------------------------------------------------------
// generators
// --------------------------------------------------------------------------
namespace {

struct EdgeTag {
  uint32_t a, b;
  int id;
};

// O-grid between an inner closed curve and the box [xb0,xb1]x[0,Ly], with
// optional structured blocks up- and downstream.  `inner(k, bx, by)` returns
// the inner-curve point hit by the ray towards box point (bx,by).
template <class Inner>
Mesh channel_with_ogrid(double Lx, double Ly, double xb0, double xb1, double h, double d_first,
                        Inner inner) {
  Mesh m;
  m.dim = 2;
  const int n = std::max(4, (int)std::lround(Ly / h));
  const int K = 4 * n;
  const double w = xb1 - xb0;
  std::vector<std::array<double, 2>> box(K), inn(K);
  for (int k = 0; k < K; ++k) {
    const int side = k / n, j = k % n;
    const double t = (double)j / n;
    if (side == 0) box[k] = {xb0 + t * w, 0.0};
    if (side == 1) box[k] = {xb1, t * Ly};
    if (side == 2) box[k] = {xb1 - t * w, Ly};
    if (side == 3) box[k] = {xb0, Ly - t * Ly};
    inn[k] = inner(box[k][0], box[k][1]);
  }
  double dmin = 1e300;
  for (int k = 0; k < K; ++k) dmin = std::min(dmin, std::hypot(box[k][0] - inn[k][0], box[k][1] - inn[k][1]));
  int nr;
  const std::vector<double> s = graded(dmin, d_first, h, &nr);
  auto og = [&](int k, int j) { return (uint32_t)(j * K + ((k % K + K) % K)); };
  for (int j = 0; j <= nr; ++j)
    for (int k = 0; k < K; ++k) {
      m.xyz.push_back(inn[k][0] + s[j] * (box[k][0] - inn[k][0]));
      m.xyz.push_back(inn[k][1] + s[j] * (box[k][1] - inn[k][1]));
    }
  // snap the outer ring exactly onto the box
  for (int k = 0; k < K; ++k) {
    m.xyz[2 * og(k, nr)] = box[k][0];
    m.xyz[2 * og(k, nr) + 1] = box[k][1];
  }
  std::vector<EdgeTag> tags;
  // upstream block
  const int nxu = xb0 > 1e-12 ? std::max(1, (int)std::lround(xb0 / h)) : 0;
  auto up = [&](int ix, int iy) -> uint32_t {  // ix = 0..nxu, iy = 0..n
    if (ix == nxu) return og((K - iy) % K, nr);
    return (uint32_t)((nr + 1) * K + ix * (n + 1) + iy);
  };
  for (int ix = 0; ix < nxu; ++ix)
    for (int iy = 0; iy <= n; ++iy) {
      m.xyz.push_back(xb0 * ix / nxu);
      m.xyz.push_back(Ly * iy / n);
    }
  const uint32_t down_base = (uint32_t)(m.xyz.size() / 2);
  const int nxd = Lx - xb1 > 1e-12 ? std::max(1, (int)std::lround((Lx - xb1) / h)) : 0;
  auto dn = [&](int ix, int iy) -> uint32_t {  // ix = 0..nxd, iy = 0..n
    if (ix == 0) return og(n + iy, nr);
    return down_base + (uint32_t)((ix - 1) * (n + 1) + iy);
  };
  for (int ix = 1; ix <= nxd; ++ix)
    for (int iy = 0; iy <= n; ++iy) {
      m.xyz.push_back(ix == nxd ? Lx : xb1 + (Lx - xb1) * ix / nxd);
      m.xyz.push_back(Ly * iy / n);
    }
  // cells: upstream, O-grid ring by ring, downstream
  for (int ix = 0; ix < nxu; ++ix)
    for (int iy = 0; iy < n; ++iy)
      push_quad(m.cells, up(ix, iy), up(ix + 1, iy), up(ix + 1, iy + 1), up(ix, iy + 1));
  for (int j = 0; j < nr; ++j)
    for (int k = 0; k < K; ++k)  // (k,j)->(k,j+1) outward, k counter-clockwise
      push_quad(m.cells, og(k, j), og(k, j + 1), og(k + 1, j + 1), og(k + 1, j));
  for (int ix = 0; ix < nxd; ++ix)
    for (int iy = 0; iy < n; ++iy)
      push_quad(m.cells, dn(ix, iy), dn(ix + 1, iy), dn(ix + 1, iy + 1), dn(ix, iy + 1));
  // boundary edges
  for (int k = 0; k < K; ++k) tags.push_back({og(k, 0), og(k + 1, 0), 4});
  for (int k = 0; k < n; ++k) tags.push_back({og(k, nr), og(k + 1, nr), 0});
  for (int k = 2 * n; k < 3 * n; ++k) tags.push_back({og(k, nr), og(k + 1, nr), 2});
  if (nxu == 0)
    for (int k = 3 * n; k < 4 * n; ++k) tags.push_back({og(k, nr), og(k + 1, nr), 3});
  if (nxd == 0)
    for (int k = n; k < 2 * n; ++k) tags.push_back({og(k, nr), og(k + 1, nr), 1});
  for (int ix = 0; ix < nxu; ++ix) {
    tags.push_back({up(ix, 0), up(ix + 1, 0), 0});
    tags.push_back({up(ix, n), up(ix + 1, n), 2});
  }
  for (int iy = 0; iy < n && nxu; ++iy) tags.push_back({up(0, iy), up(0, iy + 1), 3});
  for (int ix = 0; ix < nxd; ++ix) {
    tags.push_back({dn(ix, 0), dn(ix + 1, 0), 0});
    tags.push_back({dn(ix, n), dn(ix + 1, n), 2});
  }
  for (int iy = 0; iy < n && nxd; ++iy) tags.push_back({dn(nxd, iy), dn(nxd, iy + 1), 1});
  for (auto &t : tags) {
    m.bfaces.push_back(t.a);
    m.bfaces.push_back(t.b);
    m.bids.push_back(t.id);
  }
  orient_cells(m);
  return m;
}

}  // namespace

Mesh gen_channel2d_circle(double Lx, double Ly, double cx, double cy, double r, double h) {
  double xb0 = cx - 0.5 * Ly;
  if (xb0 < 2.0 * h) xb0 = 0.0;
  double xb1 = xb0 + Ly;
  if (Lx - xb1 < 2.0 * h) xb1 = Lx;
  const int n = std::max(4, (int)std::lround(Ly / h));
  const double arc = 2.0 * M_PI * r / (4 * n);
  return channel_with_ogrid(Lx, Ly, xb0, xb1, h, std::max(1.5 * arc, 0.25 * h), [&](double bx, double by) {
    const double dx = bx - cx, dy = by - cy, l = std::hypot(dx, dy);
    return std::array<double, 2>{cx + r * dx / l, cy + r * dy / l};
  });
}

Mesh gen_channel2d_plain(double Lx, double Ly, int nx, int ny) {
  Mesh m;
  m.dim = 2;
  auto id = [&](int i, int j) { return (uint32_t)(i * (ny + 1) + j); };
  for (int i = 0; i <= nx; ++i)
    for (int j = 0; j <= ny; ++j) {
      m.xyz.push_back(Lx * i / nx);
      m.xyz.push_back(Ly * j / ny);
    }
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) push_quad(m.cells, id(i, j), id(i + 1, j), id(i + 1, j + 1), id(i, j + 1));
  auto tag = [&](uint32_t a, uint32_t b, int t) {
    m.bfaces.push_back(a);
    m.bfaces.push_back(b);
    m.bids.push_back(t);
  };
  for (int i = 0; i < nx; ++i) {
    tag(id(i, 0), id(i + 1, 0), 0);
    tag(id(i, ny), id(i + 1, ny), 2);
  }
  for (int j = 0; j < ny; ++j) {
    tag(id(0, j), id(0, j + 1), 3);
    tag(id(nx, j), id(nx, j + 1), 1);
  }
  orient_cells(m);
  return m;
}

Mesh gen_channel2d_square(double Lx, double Ly, double ox, double oy, double s, double h) {
  auto lines = [&](std::vector<double> brk) {
    std::vector<double> g{brk[0]};
    for (size_t k = 0; k + 1 < brk.size(); ++k) {
      const double len = brk[k + 1] - brk[k];
      const int n = std::max(1, (int)std::lround(len / h));
      for (int i = 1; i <= n; ++i) g.push_back(i == n ? brk[k + 1] : brk[k] + len * i / n);
    }
    return g;
  };
  const std::vector<double> gx = lines({0.0, ox, ox + s, Lx}), gy = lines({0.0, oy, oy + s, Ly});
  const int nx = (int)gx.size() - 1, ny = (int)gy.size() - 1;
  int hx0 = 0, hx1 = 0, hy0 = 0, hy1 = 0;
  for (int i = 0; i <= nx; ++i) {
    if (gx[i] == ox) hx0 = i;
    if (gx[i] == ox + s) hx1 = i;
  }
  for (int j = 0; j <= ny; ++j) {
    if (gy[j] == oy) hy0 = j;
    if (gy[j] == oy + s) hy1 = j;
  }
  Mesh m;
  m.dim = 2;
  auto id = [&](int i, int j) { return (uint32_t)(i * (ny + 1) + j); };
  for (int i = 0; i <= nx; ++i)
    for (int j = 0; j <= ny; ++j) {
      m.xyz.push_back(gx[i]);
      m.xyz.push_back(gy[j]);
    }
  auto in_hole = [&](int i, int j) { return i >= hx0 && i < hx1 && j >= hy0 && j < hy1; };
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j)
      if (!in_hole(i, j)) push_quad(m.cells, id(i, j), id(i + 1, j), id(i + 1, j + 1), id(i, j + 1));
  auto tag = [&](uint32_t a, uint32_t b, int t) {
    m.bfaces.push_back(a);
    m.bfaces.push_back(b);
    m.bids.push_back(t);
  };
  for (int i = 0; i < nx; ++i) {
    tag(id(i, 0), id(i + 1, 0), 0);
    tag(id(i, ny), id(i + 1, ny), 2);
  }
  for (int j = 0; j < ny; ++j) {
    tag(id(0, j), id(0, j + 1), 3);
    tag(id(nx, j), id(nx, j + 1), 1);
  }
  for (int i = hx0; i < hx1; ++i) {
    tag(id(i, hy0), id(i + 1, hy0), 4);
    tag(id(i, hy1), id(i + 1, hy1), 4);
  }
  for (int j = hy0; j < hy1; ++j) {
    tag(id(hx0, j), id(hx0, j + 1), 4);
    tag(id(hx1, j), id(hx1, j + 1), 4);
  }
  drop_unused_vertices(m);
  orient_cells(m);
  return m;
}

// Airfoil pre-processing of the reference (mesh/test.py:25-41, 155-168; tests/2D/test_naca/run_test.sh:7-9):
// contour points with the chord on [0,1] are shifted to mid-chord (x - 0.5), scaled to `chord` and turned
// clockwise by the angle of attack (rotate(+angle) in test.py multiplies by R(-angle)); then placed at (cx, cy).
std::vector<std::array<double, 2>> place_airfoil(const std::vector<std::array<double, 2>> &unit, double chord,
                                                 double aoa_deg, double cx, double cy) {
  const double a = -aoa_deg * M_PI / 180.0;
  std::vector<std::array<double, 2>> pts;
  for (const auto &p : unit) {
    const double X = (p[0] - 0.5) * chord, Y = p[1] * chord;
    pts.push_back({cx + std::cos(a) * X - std::sin(a) * Y, cy + std::sin(a) * X + std::cos(a) * Y});
  }
  return pts;
}

// mesh/naca.dat / mesh/naca2412.dat layout (mesh/test.py:8-21): a name line, then "x y" pairs from the trailing
// edge over the upper side to the leading edge and back along the lower side.
std::vector<std::array<double, 2>> read_airfoil_dat(const std::string &path, std::string *name) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("read_airfoil_dat: cannot open " + path);
  std::string line;
  std::getline(f, line);
  if (name) *name = line;
  std::vector<std::array<double, 2>> pts;
  double x, y;
  while (f >> x >> y) pts.push_back({x, y});
  if (pts.size() < 8) throw std::runtime_error("read_airfoil_dat: fewer than 8 contour points in " + path);
  // a closing point that repeats the first one would give a zero-length segment
  if (std::hypot(pts.front()[0] - pts.back()[0], pts.front()[1] - pts.back()[1]) < 1e-12) pts.pop_back();
  return pts;
}

std::vector<std::array<double, 2>> naca4_contour(int naca4, int n_around) {
  // NACA 4-digit thickness/camber (closed trailing edge variant so that the
  // contour is a closed curve; mesh/naca2412.dat is the same family sampled
  // at 35 points with a 0.0026 blunt edge).
  const double mc = (naca4 / 1000) / 100.0, pc = ((naca4 / 100) % 10) / 10.0, tc = (naca4 % 100) / 100.0;
  auto yt = [&](double x) {
    return 5.0 * tc * (0.2969 * std::sqrt(x) - 0.1260 * x - 0.3516 * x * x + 0.2843 * x * x * x - 0.1036 * x * x * x * x);
  };
  auto camber = [&](double x, double *dy) {
    if (pc <= 0) {
      *dy = 0;
      return 0.0;
    }
    if (x < pc) {
      *dy = 2 * mc / (pc * pc) * (pc - x);
      return mc / (pc * pc) * (2 * pc * x - x * x);
    }
    *dy = 2 * mc / ((1 - pc) * (1 - pc)) * (pc - x);
    return mc / ((1 - pc) * (1 - pc)) * (1 - 2 * pc + 2 * pc * x - x * x);
  };
  // counter-clockwise: trailing edge -> upper side -> leading edge -> lower side
  const int half = std::max(8, n_around / 2);
  std::vector<std::array<double, 2>> pts;
  for (int i = 0; i < 2 * half; ++i) {
    const bool upper = i < half;
    const double beta = M_PI * (upper ? i : (2 * half - i)) / half;  // 0 at TE, pi at LE
    const double x = 0.5 * (1 + std::cos(beta));
    double dy;
    const double yc = camber(x, &dy), th = std::atan(dy), t = yt(x);
    double px = upper ? x - t * std::sin(th) : x + t * std::sin(th);
    double py = upper ? yc + t * std::cos(th) : yc - t * std::cos(th);
    if (i == 0) {
      px = 1.0;
      py = 0.0;
    }
    pts.push_back({px, py});
  }
  return pts;
}

Mesh gen_naca2d(double Lx, double Ly, double cx, double cy, int naca4, double aoa_deg, double chord,
                int n_around, int n_radial) {
  return gen_airfoil2d(Lx, Ly, cx, cy, place_airfoil(naca4_contour(naca4, n_around), chord, aoa_deg, cx, cy), chord,
                       n_radial);
}

// O-grid between a closed contour around (cx, cy) and the box [0,Lx] x [0,Ly]; ids: contour 4, box sides
// 0 bottom, 1 right (outlet), 2 top, 3 left (inlet) as in mesh/NACA_2412.geo:108-113 and mesh/test.py's writer.
Mesh gen_airfoil2d(double Lx, double Ly, double cx, double cy, const std::vector<std::array<double, 2>> &pts,
                   double chord, int n_radial) {
  // rays from the centre (cx,cy): the contour is star-shaped with respect to
  // its mid-chord point for the cambers/thicknesses of the 4-digit family.
  struct Ray {
    double th;
    std::array<double, 2> in;
  };
  std::vector<Ray> rays;
  for (auto &p : pts) rays.push_back({std::atan2(p[1] - cy, p[0] - cx), p});
  auto hit_contour = [&](double th) {
    const double dx = std::cos(th), dy = std::sin(th);
    for (size_t i = 0; i < pts.size(); ++i) {
      const auto &a = pts[i], &b = pts[(i + 1) % pts.size()];
      const double ex = b[0] - a[0], ey = b[1] - a[1];
      const double den = dx * ey - dy * ex;
      if (std::fabs(den) < 1e-300) continue;
      const double t = ((a[0] - cx) * ey - (a[1] - cy) * ex) / den;
      const double u = ((a[0] - cx) * dy - (a[1] - cy) * dx) / den;
      if (t > 0 && u >= 0 && u <= 1) return std::array<double, 2>{cx + t * dx, cy + t * dy};
    }
    throw std::runtime_error("gen_naca2d: ray misses the contour");
  };
  const double corners[4][2] = {{0, 0}, {Lx, 0}, {Lx, Ly}, {0, Ly}};
  for (auto &c : corners) {
    const double th = std::atan2(c[1] - cy, c[0] - cx);
    rays.push_back({th, hit_contour(th)});
  }
  std::sort(rays.begin(), rays.end(), [](const Ray &a, const Ray &b) { return a.th < b.th; });
  // drop rays that coincide in angle
  std::vector<Ray> uniq;
  for (auto &r : rays)
    if (uniq.empty() || r.th - uniq.back().th > 1e-9) uniq.push_back(r);
  rays.swap(uniq);
  const int K = (int)rays.size();
  auto hit_box = [&](double th, int *side) {
    const double dx = std::cos(th), dy = std::sin(th);
    double best = 1e300;
    const double cand[4] = {dy < 0 ? (0 - cy) / dy : 1e300, dx > 0 ? (Lx - cx) / dx : 1e300,
                            dy > 0 ? (Ly - cy) / dy : 1e300, dx < 0 ? (0 - cx) / dx : 1e300};
    for (int s = 0; s < 4; ++s)
      if (cand[s] < best) {
        best = cand[s];
        *side = s;
      }
    return std::array<double, 2>{cx + best * dx, cy + best * dy};
  };
  Mesh m;
  m.dim = 2;
  int nr;
  const double dfar = std::min(std::min(cx, Lx - cx), std::min(cy, Ly - cy));
  std::vector<double> s = graded(1.0, 0.02 * chord / dfar, 1.0 / std::max(4, n_radial / 4), &nr);
  (void)n_radial;
  std::vector<int> side(K);
  for (int j = 0; j <= nr; ++j)
    for (int k = 0; k < K; ++k) {
      const auto out = hit_box(rays[k].th, &side[k]);
      m.xyz.push_back(rays[k].in[0] + s[j] * (out[0] - rays[k].in[0]));
      m.xyz.push_back(rays[k].in[1] + s[j] * (out[1] - rays[k].in[1]));
    }
  auto og = [&](int k, int j) { return (uint32_t)(j * K + ((k % K + K) % K)); };
  for (int j = 0; j < nr; ++j)
    for (int k = 0; k < K; ++k) push_quad(m.cells, og(k, j), og(k, j + 1), og(k + 1, j + 1), og(k + 1, j));
  for (int k = 0; k < K; ++k) {
    m.bfaces.insert(m.bfaces.end(), {og(k, 0), og(k + 1, 0)});
    m.bids.push_back(4);
    // outer edge k..k+1 lies on the side of its midpoint angle
    int sd;
    double th2 = rays[(k + 1) % K].th;
    if (k + 1 == K) th2 += 2 * M_PI;
    hit_box(0.5 * (rays[k].th + th2), &sd);
    m.bfaces.insert(m.bfaces.end(), {og(k, nr), og(k + 1, nr)});
    m.bids.push_back(sd);  // 0 bottom, 1 right(outlet), 2 top, 3 left(inlet)
  }
  orient_cells(m);
  return m;
}

Mesh extrude_to_tets(const Mesh &m2, double Lz, int nz) {
  if (m2.dim != 2) throw std::runtime_error("extrude_to_tets: need a 2D mesh");
  Mesh m;
  m.dim = 3;
  const uint32_t nv2 = (uint32_t)m2.n_verts();
  m.xyz.reserve((size_t)nv2 * (nz + 1) * 3);
  for (int l = 0; l <= nz; ++l)
    for (uint32_t v = 0; v < nv2; ++v) {
      m.xyz.push_back(m2.xyz[2 * v]);
      m.xyz.push_back(m2.xyz[2 * v + 1]);
      m.xyz.push_back(l == nz ? Lz : Lz * l / nz);
    }
  // prism vertex rotations bringing the smallest index to the front
  static const int rot[6][6] = {{0, 1, 2, 3, 4, 5}, {1, 2, 0, 4, 5, 3}, {2, 0, 1, 5, 3, 4},
                                {3, 5, 4, 0, 2, 1}, {4, 3, 5, 1, 0, 2}, {5, 4, 3, 2, 1, 0}};
  m.cells.reserve(m2.n_cells() * nz * 12);
  for (int l = 0; l < nz; ++l)
    for (size_t t = 0; t < m2.n_cells(); ++t) {
      uint32_t P[6];
      for (int k = 0; k < 3; ++k) {
        P[k] = l * nv2 + m2.cells[3 * t + k];
        P[k + 3] = (l + 1) * nv2 + m2.cells[3 * t + k];
      }
      int s = 0;
      for (int k = 1; k < 6; ++k)
        if (P[k] < P[s]) s = k;
      uint32_t V[6];
      for (int k = 0; k < 6; ++k) V[k] = P[rot[s][k]];
      if (std::min(V[1], V[5]) < std::min(V[2], V[4]))
        m.cells.insert(m.cells.end(), {V[0], V[1], V[2], V[5], V[0], V[1], V[5], V[4], V[0], V[4], V[5], V[3]});
      else
        m.cells.insert(m.cells.end(), {V[0], V[1], V[2], V[4], V[0], V[4], V[2], V[5], V[0], V[4], V[5], V[3]});
    }
  auto tri = [&](uint32_t a, uint32_t b, uint32_t c, int id) {
    m.bfaces.insert(m.bfaces.end(), {a, b, c});
    m.bids.push_back(id);
  };
  for (size_t t = 0; t < m2.n_cells(); ++t) {
    const uint32_t *v = &m2.cells[3 * t];
    tri(v[0], v[1], v[2], 0);
    tri(nz * nv2 + v[0], nz * nv2 + v[1], nz * nv2 + v[2], 0);
  }
  static const int map3[5] = {2, 1, 2, 3, 4};
  for (size_t e = 0; e < m2.n_bfaces(); ++e) {
    const uint32_t a = m2.bfaces[2 * e], b = m2.bfaces[2 * e + 1];
    const int id2 = m2.bids[e];
    const int id = (id2 >= 0 && id2 < 5) ? map3[id2] : id2;
    for (int l = 0; l < nz; ++l) {
      const uint32_t a0 = l * nv2 + a, b0 = l * nv2 + b, a1 = a0 + nv2, b1 = b0 + nv2;
      if (a < b) {  // diagonal a0-b1 (a0 is the smallest index of the quad)
        tri(a0, b0, b1, id);
        tri(a0, b1, a1, id);
      } else {  // diagonal b0-a1
        tri(a0, b0, a1, id);
        tri(b0, b1, a1, id);
      }
    }
  }
  orient_cells(m);
  return m;
}

Mesh gen_named(const std::string &name, double h) {
  if (name == "2d-cylinder")  // mesh/domain2D.geo:2-10
    return gen_channel2d_circle(2.2, 0.41, 0.2, 0.2, 0.05, h);
  if (name == "3d-square") {  // mesh/domain3D.geo:2-12
    Mesh m2 = gen_channel2d_square(2.5, 0.41, 0.45, 0.15, 0.1, h);
    return extrude_to_tets(m2, 0.41, std::max(1, (int)std::lround(0.41 / h)));
  }
  if (name == "3d-cylinder") {  // mesh/domain3D2.geo:2-9
    Mesh m2 = gen_channel2d_circle(2.5, 0.41, 0.45, 0.20, 0.05, h);
    return extrude_to_tets(m2, 0.41, std::max(1, (int)std::lround(0.41 / h)));
  }
  if (name == "naca2412") {  // mesh/NACA_2412.geo:2-9
    const int n_around = std::max(32, (int)std::lround(2.1 / h));
    return gen_naca2d(35.0, 20.0, 10.0, 10.0, 2412, 0.0, 1.0, n_around, 48);
  }
  if (name == "naca2408-run_test") {  // tests/2D/test_naca/run_test.sh:7-9 at angle 0: chord 0.4 in the 2.2 x 1.0 box
    const int n_around = std::max(32, (int)std::lround(0.85 / h));
    return gen_naca2d(2.2, 1.0, 0.4, 0.5, 2408, 0.0, 0.4, n_around, 48);
  }
  if (name == "channel2d") {
    const int ny = std::max(2, (int)std::lround(0.41 / h));
    return gen_channel2d_plain(2.2, 0.41, std::max(2, (int)std::lround(2.2 / h)), ny);
  }
  if (name == "channel3d") {
    const int ny = std::max(2, (int)std::lround(0.41 / h));
    Mesh m2 = gen_channel2d_plain(2.5, 0.41, std::max(2, (int)std::lround(2.5 / h)), ny);
    return extrude_to_tets(m2, 0.41, ny);
  }
  throw std::runtime_error("gen_named: unknown mesh '" + name + "'");
}

}  // namespace nsb
