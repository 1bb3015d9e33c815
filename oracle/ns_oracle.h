/* CPU oracle for the per-time-step hot path of
 * denisuntila/NavierStokes-Capoferri_Cecchettini_Untila.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path may include, link or
 * call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker and as the CPU baseline.
 *
 * PARITY UNPINNED: the reference's arithmetic lives in un-vendored deal.II /
 * Trilinos (absent from this image and from /root/reference) and the
 * reference's tests hold no golden vectors.  This restatement
 * follows reference src/NavierStokes.cpp line by line for the application code
 * and SURVEY.md Appendix A for the library semantics; it is pinned only by its
 * own analytic known-answer tests (tests/test_oracle_kat.py).
 */
#ifndef NS_ORACLE_H
#define NS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nso nso;

/* Quadrature tables. */
enum { NSO_QUAD_DEALII93 = 0, NSO_QUAD_DEALII95 = 1 };
/* Inlet profile kinds (drivers: tests/2D/test_01/src/test_01.cpp:18-42 etc.). */
enum { NSO_INLET_PARABOLIC = 0, NSO_INLET_UNIFORM = 1 };

/* Builds topology, deal.II-compatible numbering and the block sparsity
 * pattern (NavierStokes::setup, reference :4-131). cells hold dim+1 vertex ids,
 * bfaces dim vertex ids with their boundary id. */
nso *nso_create(int dim, int64_t n_verts, const double *xyz, int64_t n_cells, const uint32_t *cells,
                int64_t n_bfaces, const uint32_t *bfaces, const int32_t *bids, int quad_rule);
void nso_destroy(nso *);

/* sizes: [0] n_u, [1] n_p, [2] nnz A00, [3] nnz A01, [4] nnz A10, [5] nnz S,
 *        [6] dofs_per_cell, [7] n_q, [8] n_q_face, [9] n_bc */
void nso_sizes(const nso *, int64_t out[10]);
void nso_get_cell_dofs(const nso *, uint32_t *out);                 /* n_cells*dpc */
/* block: 0 A00, 1 A01, 2 A10, 3 S.  Block-local column indices. */
void nso_get_pattern(const nso *, int block, int64_t *rowptr, uint32_t *colind);
void nso_get_values(const nso *, int block, double *vals);
void nso_get_rhs(const nso *, double *out);                         /* n_u+n_p */
void nso_get_lumped(const nso *, double *out);                      /* deltat_lumped_mass_inv */
void nso_get_bc(const nso *, uint32_t *dofs, double *values);       /* after assemble */

void nso_set_params(nso *, double deltat, double nu);
/* bc_diag_mode: 0 keep the assembled diagonal of constrained rows (upstream
 * TrilinosWrappers::SparseMatrix::clear_row), 1 overwrite with the first
 * non-zero diagonal entry (SURVEY.md A.7 reading). */
void nso_set_bc_diag_mode(nso *, int mode);
void nso_set_inlet(nso *, int kind, double U_m, double H, int time_sin);
double nso_mean_velocity(const nso *, double time);
/* nu = U*Diameter/Re with U = get_mean_vel() at the current inlet time
 * (reference :332-341; Diameter = 0.4, NavierStokes.hpp:256). */
double nso_set_re_number(nso *, int Re);
void nso_set_solution(nso *, const double *x);  /* solution_owned and ghosted solution */
void nso_get_solution(const nso *, double *x);
/* outer GMRES: tolerance factor (1e-6), max_n_tmp_vectors (30), max its (10000);
 * inner tolerance factor (1e-2). */
void nso_set_solver(nso *, double outer_rtol, int n_tmp_vectors, int max_it, double inner_rtol);

void nso_assemble(nso *, double time);                       /* reference :133-330 */
/* reference :344-397 with PreconditionASIMPLE :934-995.  Returns 0, or 1 when
 * the outer iteration count limit is hit (deal.II would throw). */
int nso_solve_time_step(nso *, int *iters, double *t_prec, double *t_solve);
void nso_compute_forces(nso *, double time, double out[4]);  /* drag, lift, cd, cl; reference :831-929 */
/* y = A x on the assembled block matrix (n_u+n_p). */
void nso_vmult(const nso *, const double *x, double *y);
/* Number of OpenMP threads used for the element-matrix batch (scatter stays
 * serial and in cell order). */
void nso_set_threads(nso *, int n);

/* ---- CPU baseline mode (bench.py only; ns_baseline.cpp) --------------------------------
 * The same algorithm the way `mpirun -n P` of the reference runs it, with P OpenMP threads:
 * cells partitioned into P subdomains (reference :19-23), every subdomain assembles its own
 * cells (:166), rank-local ILU(0) of the diagonal blocks (Ifpack, overlap 0; :958-959), all
 * SpMV / dot / axpy of the outer and inner GMRES solves on all threads.  cell_part[c] in
 * [0, n_parts).  The checker functions above are unaffected. */
int nso_baseline_partition(nso *, int n_parts, const int32_t *cell_part);
void nso_baseline_assemble(nso *, double time);
int nso_baseline_solve_time_step(nso *, int *iters, double *t_prec, double *t_solve);

#ifdef __cplusplus
}
#endif
#endif
