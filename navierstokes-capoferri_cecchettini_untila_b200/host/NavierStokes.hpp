// Drop-in host facade: the public interface of the reference's NavierStokes
// class (reference src/NavierStokes.hpp:51-271) on top of the B200 hot path.
//
// Same customisation points as the reference: compile with -DDIM=2|3; define
// NS_INPUT before including this header to supply
// InletVelocity::{vector_value, value, get_mean_vel} in the driver
// (reference NavierStokes.hpp:77-121).  The reference's drivers
// (tests/*/src/*.cpp) compile unmodified against this header: the few deal.II
// names they touch (Point, Vector, Function, Utilities::MPI::MPI_InitFinalize)
// are provided below as minimal stand-ins.
//
// Everything numerical is done by libnsb.so through include/nsb.h; mesh,
// numbering and sparsity pattern come from libnsb_host.so.  There is no CPU
// fallback: setup() throws when no CUDA device is usable.
#ifndef NSB_NAVIER_STOKES_FACADE_HPP
#define NSB_NAVIER_STOKES_FACADE_HPP

#include <array>
#include <cmath>
#include <cstddef>
#include <fstream>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef DIM
#error "compile with -DDIM=2 or -DDIM=3 (reference tests/*/common/cmake-common.cmake:4)"
#endif

// ---- minimal stand-ins for the deal.II names used by the drivers -----------
template <int dim_>
class Point {
public:
  Point() { c.fill(0.0); }
  double &operator[](unsigned int i) { return c[i]; }
  const double &operator[](unsigned int i) const { return c[i]; }

private:
  std::array<double, dim_> c;
};

template <typename Number>
class Vector {
public:
  Vector() = default;
  explicit Vector(std::size_t n) : v(n, Number(0)) {}
  Number &operator[](std::size_t i) { return v[i]; }
  const Number &operator[](std::size_t i) const { return v[i]; }
  std::size_t size() const { return v.size(); }

private:
  std::vector<Number> v;
};

template <int dim_>
class Function {
public:
  explicit Function(unsigned int n_components_ = 1) : n_components(n_components_) {}
  virtual ~Function() = default;
  virtual double value(const Point<dim_> &, const unsigned int = 0) const { return 0.0; }
  virtual void vector_value(const Point<dim_> &p, Vector<double> &values) const {
    for (unsigned int i = 0; i < n_components; ++i) values[i] = value(p, i);
  }
  void set_time(double t) { time_ = t; }
  double get_time() const { return time_; }
  const unsigned int n_components;

private:
  double time_ = 0.0;
};

namespace Utilities {
namespace MPI {
// One process per GPU: rank/size come from the launcher's environment
// (RANK / WORLD_SIZE, as set by torchrun or mpirun wrappers); nothing to
// initialise for a single process.
struct MPI_InitFinalize {
  MPI_InitFinalize(int &, char **&) {}
};
}  // namespace MPI
}  // namespace Utilities

struct nsb_ctx;
namespace nsb {
struct Problem;
struct LocalProblem;
}

class NavierStokes {
public:
  static constexpr unsigned int dim = DIM;

  class ForcingTerm : public Function<dim> {
  public:
    double value(const Point<dim> &, const unsigned int = 0) const override { return 0.0; }
  };

  class InletVelocity : public Function<dim> {
  public:
    InletVelocity() : Function<dim>(dim + 1) {}
#ifndef NS_INPUT
    // defaults of the reference when the driver does not define NS_INPUT
    void vector_value(const Point<dim> &, Vector<double> &values) const override {
      for (unsigned int i = 0; i < dim + 1; ++i) values[i] = 0.0;
      values[0] = 3.0;
    }
    double value(const Point<dim> &, const unsigned int component = 0) const override {
      return component == 0 ? 3.0 : 0.0;
    }
    double get_mean_vel() { return 2.0 / 3.0; }
#else
    void vector_value(const Point<dim> &p, Vector<double> &values) const override;
    double value(const Point<dim> &p, const unsigned int component = 0) const override;
    double get_mean_vel();
#endif
  };

  class InitialConditions : public Function<dim> {
  public:
    InitialConditions() : Function<dim>(dim + 1) {}
    double value(const Point<dim> &, const unsigned int = 0) const override { return 0.0; }
  };

  NavierStokes(const std::string &mesh_file_name_, const unsigned int &degree_velocity_,
               const unsigned int &degree_pressure_, const double &deltat_, const double &T_,
               const unsigned int &step_);
  ~NavierStokes();
  NavierStokes(const NavierStokes &) = delete;
  NavierStokes &operator=(const NavierStokes &) = delete;

  void setup();
  void set_re_number(int Re);
  void solve_time_step(std::ostream &oss);
  void assemble(const double &time);
  void output(const unsigned int &time_step) const;
  void solve(unsigned int time_step = 0);
  void export_data(const unsigned int &time_step);
  void compute_ordered_dofs_indices();
  void import_data(const unsigned int &time_step);
  void post_process(const unsigned int &initial_time_step, const unsigned int &final_time_step,
                    const unsigned int &step);
  void compute_forces(const double &time);

  // ---- additions (not in the reference) ---------------------------------
  // Quadrature table family (deal.II 9.3.x vs >= 9.4); default >= 9.4.
  void set_quadrature_rule(int rule_id) { quad_rule = rule_id; }
  // Outer GMRES and inner-sweep knobs; defaults reproduce the reference's
  // stopping rule (1e-6 ||rhs||, restart 28, max 10000, alpha 0.5).
  void set_solver_options(double gmres_rtol, int restart, int max_it, int sweeps_F, int sweeps_S);
  const std::vector<double> &get_solution() const { return solution; }
  double get_drag() const { return drag; }
  double get_lift() const { return lift; }
  double get_cd() const { return cd; }
  double get_cl() const { return cl; }
  unsigned int last_gmres_iterations() const { return last_iters; }

protected:
  const std::string mesh_file_name;
  const unsigned int mpi_size, mpi_rank;
  const unsigned int degree_velocity, degree_pressure;

  std::unique_ptr<nsb::Problem> problem;  // mesh, dof numbering, patterns (host)
  std::unique_ptr<nsb::LocalProblem> local;  // this rank's part when mpi_size > 1 (one process per GPU)
  std::vector<double> local_vec;             // [u owned | u ghost | p] staging for the distributed context
  nsb_ctx *ctx = nullptr;                 // device-resident system (system_matrix, system_rhs, solution_owned)
  std::vector<double> solution;           // host mirror of the ghosted solution (reference NavierStokes.hpp:252)
  std::vector<unsigned int> renumbered_dofs;
  std::vector<double> bc_values;

  double nu = 1.e-3;            // NavierStokes.hpp:254
  const double p_out = 0.0;     // NavierStokes.hpp:255
  const double Diameter = 0.4;  // NavierStokes.hpp:256
  const double deltat;
  double time = 0.0;
  const double T;
  const unsigned int step;

  ForcingTerm forcing_term;
  InletVelocity inlet_velocity;
  InitialConditions initial_conditions;

  double drag = 0, lift = 0, cd = 0, cl = 0;
  unsigned int last_iters = 0;
  int quad_rule = 1;
  double opt_rtol = 1e-6;
  int opt_restart = 28, opt_max_it = 10000, opt_sweeps_F = 0, opt_sweeps_S = 0;

  void check(int rc, const char *what) const;
  void refresh_dirichlet(double time);
  void setup_distributed();
  void push_solution();  // host `solution` (canonical numbering) -> device
  void pull_solution();  // device -> host `solution` (collective when mpi_size > 1)
};

#endif
