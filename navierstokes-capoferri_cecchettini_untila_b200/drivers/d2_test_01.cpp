// Driver for the reference case d2_test_01 (parameters: SURVEY.md §4).
#define NS_INPUT
#include "NavierStokes.hpp"
#include "driver_common.hpp"

static constexpr double U_m = 0.3;
static constexpr double H = 0.41;

double NavierStokes::InletVelocity::value(const Point<dim> &p, const unsigned int component) const {
  (void)p;
  return component == 0 ? (4 * U_m * p[1] * (H - p[1]) / (H * H)) : 0.0;
}
void NavierStokes::InletVelocity::vector_value(const Point<dim> &p, Vector<double> &values) const {
  for (unsigned int i = 0; i < dim + 1; ++i) values[i] = value(p, i);
}
double NavierStokes::InletVelocity::get_mean_vel() { return 2.0 * U_m / 3.0; }

int main(int argc, char **argv) { return run_case(argc, argv, "d2_test_010", 0.01, 2.0, 10, 20); }
