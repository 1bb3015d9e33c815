// Driver for the reference case d2_test_naca (parameters: SURVEY.md §4).
#define NS_INPUT
#include "NavierStokes.hpp"
#include "driver_common.hpp"

static constexpr double U_m = 1.0;
static constexpr double H = 0.41;

double NavierStokes::InletVelocity::value(const Point<dim> &p, const unsigned int component) const {
  (void)p;
  return component == 0 ? (U_m) : 0.0;
}
void NavierStokes::InletVelocity::vector_value(const Point<dim> &p, Vector<double> &values) const {
  for (unsigned int i = 0; i < dim + 1; ++i) values[i] = value(p, i);
}
double NavierStokes::InletVelocity::get_mean_vel() { return U_m; }

int main(int argc, char **argv) { return run_case(argc, argv, "d2_test_naca0", 0.01, 1.0, 2, 0); }
