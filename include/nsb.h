/* nsb: C ABI of the B200 (sm_100a) hot path -- per-time-step assembly of the
 * semi-implicit Oseen system on Taylor-Hood P2/P1 simplices, the aSIMPLE-
 * preconditioned GMRES solve, and the Cd/Cl force integrals.
 *
 * This is the boundary a NavierStokes-style host class binds to (SURVEY.md
 * §8b).  Plain pointers and sizes only; every function returns 0 on success or
 * a negative NSB_E* code and never throws; nsb_last_error(ctx) explains.  Host
 * pointers are not retained after a call returns.  There is no CPU fallback:
 * nsb_create fails when no CUDA device is usable.
 *
 * Reference interface each entry point replaces (file:line in
 * /root/reference/src):
 *   nsb_set_mesh / nsb_set_dofs / nsb_set_pattern   NavierStokes::setup            NavierStokes.cpp:4-131
 *   nsb_set_quadrature                              QGaussSimplex(fe->degree+1)    NavierStokes.cpp:50,54
 *   nsb_set_params                                  ctor deltat, set_re_number nu  NavierStokes.hpp:173-189, .cpp:332-341
 *   nsb_set_solution / nsb_get_solution             solution_owned / solution      NavierStokes.hpp:251-252, .cpp:395,465
 *   nsb_set_dirichlet                               interpolate_boundary_values    NavierStokes.cpp:297-324
 *   nsb_assemble                                    NavierStokes::assemble         NavierStokes.cpp:133-330
 *   nsb_solve_time_step                             NavierStokes::solve_time_step  NavierStokes.cpp:344-397
 *                                                   + PreconditionASIMPLE          NavierStokes.cpp:934-995
 *   nsb_set_force_faces / nsb_compute_forces        NavierStokes::compute_forces   NavierStokes.cpp:831-929
 *   nsb_get_matrix_values / nsb_get_rhs / nsb_vmult parity taps on system_matrix / system_rhs (NavierStokes.hpp:248-250)
 */
#ifndef NSB_H
#define NSB_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nsb_ctx nsb_ctx;

enum {
  NSB_OK = 0,
  NSB_ECUDA = -1,     /* CUDA runtime error / no device */
  NSB_EARG = -2,      /* bad argument or call order */
  NSB_ESTRUCT = -3,   /* dofs/pattern do not have the Taylor-Hood structure */
  NSB_ENOCONV = -4,   /* GMRES hit max_it (deal.II: SolverControl::NoConvergence) */
  NSB_ENCCL = -5      /* NCCL error */
};

/* blocks of the 2x2 system matrix (A11 is empty) and the Schur approximation */
enum { NSB_A00 = 0, NSB_A01 = 1, NSB_A10 = 2, NSB_S = 3 };
/* quadrature tables: deal.II 9.3.x (2D 7-pt hard-coded order, 3D 10-pt degree 3)
 * or >= 9.4 (Witherden-Vincent: 2D 7-pt, 3D 14-pt degree 5) */
enum { NSB_QUAD_DEALII93 = 0, NSB_QUAD_DEALII95 = 1 };
/* treatment of the diagonal of constrained rows in apply_boundary_values */
enum { NSB_BCDIAG_KEEP = 0, NSB_BCDIAG_FIRST = 1 };
/* preconditioner selection (reference NavierStokes.cpp:352-373): PreconditionASIMPLE (:934-995, the one the
 * reference enables), PreconditionIdentity (NavierStokes.hpp:274-287), PreconditionAYosida (:998-1051) */
enum { NSB_PREC_ASIMPLE = 0, NSB_PREC_IDENTITY = 1, NSB_PREC_AYOSIDA = 2 };

const char *nsb_last_error(const nsb_ctx *ctx);
/* Number of CUDA devices visible (0 when none / no driver). */
int nsb_device_count(void);

int nsb_create(int dim, int device_id, nsb_ctx **out);
void nsb_destroy(nsb_ctx *ctx);

/* ---- immutable inputs (once) ------------------------------------------ */
int nsb_set_mesh(nsb_ctx *ctx, int64_t n_verts, const double *xyz, int64_t n_cells, const uint32_t *cell_verts);
/* cell_dofs: n_cells x dofs_per_cell in deal.II FESystem cell order
 * (per vertex u_0..u_{dim-1}, p; then per line u_0..u_{dim-1}); velocity dofs
 * must be dim*node+c and pressure dofs n_u+vertex. */
int nsb_set_dofs(nsb_ctx *ctx, uint32_t n_u, uint32_t n_p, const uint32_t *cell_dofs);
/* Canonical CSR of a block (rows ascending, columns ascending, block-local
 * column indices).  Blocks A00, A01, A10 are required; NSB_S is optional (the
 * structural product A10*A01 is formed on the device when absent). */
int nsb_set_pattern(nsb_ctx *ctx, int block, int64_t n_rows, const int64_t *rowptr, const uint32_t *colind);
/* Alternative to nsb_set_pattern(NSB_A00): node-level adjacency, expanded on
 * the device to the canonical A00 = nodes (x) ones(dim,dim) pattern. */
int nsb_set_node_pattern(nsb_ctx *ctx, int64_t n_nodes, const int64_t *rowptr, const uint32_t *colind);
int nsb_set_quadrature(nsb_ctx *ctx, int rule_id);
/* Finishes setup: scatter maps, diagonal positions, work vectors. Called
 * implicitly by the first nsb_assemble. */
int nsb_finalize_setup(nsb_ctx *ctx);

/* ---- parameters ------------------------------------------------------- */
int nsb_set_params(nsb_ctx *ctx, double deltat, double nu);
int nsb_set_bc_diag_mode(nsb_ctx *ctx, int mode);
/* Outer GMRES: stop when the preconditioned residual <= rtol*||rhs||_2
 * (reference :348), restart length (deal.II default 28), max iterations
 * (10000).  alpha: aSIMPLE relaxation (NavierStokes.hpp:306, 0.5). */
int nsb_set_solver(nsb_ctx *ctx, double gmres_rtol, int restart, int max_it, double alpha, int preconditioner);
/* Inner solves of aSIMPLE: fixed-degree Chebyshev-Jacobi polynomials in place
 * of the reference's ILU-preconditioned GMRES to 1e-2 (north star: Jacobi-type
 * inner sweeps).  sweeps = polynomial degree; eig_ratio = assumed
 * lambda_max/lambda_min of D^-1 M targeted by the polynomial.  sweeps <= 0
 * (the default) selects an automatic choice from the problem size. */
int nsb_set_inner(nsb_ctx *ctx, int sweeps_F, double eig_ratio_F, int sweeps_S, double eig_ratio_S);
/* Solver for the Schur block S inside aSIMPLE.  mode 0: the single-level
 * Chebyshev-Jacobi polynomial configured by nsb_set_inner.  mode 1 (default):
 * one V-cycle over an aggregation hierarchy whose smoother is the same
 * Chebyshev-Jacobi sweep (smoother_sweeps per side, default 1 -- measured on B200 at 9.7 M DoFs:
 * 485 ms/step with 1, 514 with 2, 564 with 3;
 * strength-of-connection threshold theta (measure and defaults: nsb_set_schur_strength); coarse-correction
 * scaling omega, default 1.5; `cycles` V-cycles per application, default 1).
 * Arguments <= 0 keep the defaults. */
int nsb_set_schur_solver(nsb_ctx *ctx, int mode, int smoother_sweeps, double strength_theta, double omega,
                         int cycles);

/* Strength-of-connection measure of the aggregation hierarchy: 0: -s_ij >= theta max_k(-s_ik) (default in 2D,
 * theta 0.35, no decay); 1: |s_ij| >= theta sqrt(s_ii s_jj) (default in 3D, theta 0.08, decay 0.5);
 * 2: |s_ij| >= theta max_k |s_ik|.  theta is multiplied by decay_per_level on every coarser level.
 * measure -1 / theta 0 / decay 0 select the defaults.  Rebuilds the hierarchy at the next step. */
int nsb_set_schur_strength(nsb_ctx *ctx, int measure, double theta, double decay_per_level);

/* ---- state ------------------------------------------------------------ */
int nsb_set_solution(nsb_ctx *ctx, const double *x_host);   /* n_u+n_p, host -> device */
int nsb_get_solution(nsb_ctx *ctx, double *x_host);         /* device -> host */
int nsb_set_dirichlet(nsb_ctx *ctx, int64_t n_bc, const uint32_t *dofs, const double *values);
/* Values only (same dof list), scaled by `factor` on the device: the
 * time-dependent inlets g(x,t) = g(x) * sin(pi t/8) of the *_03 drivers. */
int nsb_scale_dirichlet(nsb_ctx *ctx, double factor);
int nsb_set_force_faces(nsb_ctx *ctx, int64_t n_faces, const uint32_t *cell, const double *normal,
                        const double *measure);

/* ---- the per-time-step path ------------------------------------------- */
int nsb_assemble(nsb_ctx *ctx, double time);
int nsb_solve_time_step(nsb_ctx *ctx, int *iters, double *t_prec, double *t_solve);
/* out = {drag, lift, cd, cl}; u_mean = InletVelocity::get_mean_vel(). */
int nsb_compute_forces(nsb_ctx *ctx, double u_mean, double out[4]);

/* ---- parity taps / bench hooks ---------------------------------------- */
int nsb_get_matrix_values(nsb_ctx *ctx, int block, double *vals_host);
int nsb_get_pattern(nsb_ctx *ctx, int block, int64_t *rowptr_host, uint32_t *colind_host);
int64_t nsb_nnz(const nsb_ctx *ctx, int block);
int nsb_get_rhs(nsb_ctx *ctx, double *rhs_host);
/* deltat_lumped_mass_inv of the reference (NavierStokes.cpp:232-236, 252, 284-290), velocity block: n_u values
 * deltat / sum_cells sum_q sum_j |phi_j . phi_i JxW| (the pressure block of the reference vector is deltat/0). */
int nsb_get_lumped_mass_inv(nsb_ctx *ctx, double *out_host);
/* y = A x with host vectors (n_u+n_p). */
int nsb_vmult(nsb_ctx *ctx, const double *x_host, double *y_host);
/* Device-resident micro-benchmarks: run `reps` launches and return the mean
 * kernel time in milliseconds measured with CUDA events on the ctx stream.
 * which: 0 block SpMV y=Ax on the canonical (reference) CSR, 1 assembly
 * (zero + cell loop + Dirichlet rows), 2 preconditioner apply, 3 S = B Di Bt,
 * 4 Chebyshev sweep on F (node-block storage), 5 block SpMV on the compressed
 * storage the solver uses, 6 Chebyshev sweep on S, 7 dst0 = vec0 - Di .* (A01 p), 8 vec1 = src1 - A10 u,
 * 9 the whole F solve of one preconditioner application, 10 the whole Schur solve (incl. its exchanges on
 * several GPUs), 11 one velocity halo exchange, 12 one all-gather of the owned pressure rows, 13 one Gram-Schmidt
 * orthogonalisation (two passes) against 14 basis vectors.
 * Add 0x100 to flush L2 between repetitions. */
int nsb_bench_kernel(nsb_ctx *ctx, int which, int reps, double *ms_mean);
/* CUDA-event bracket on the context's stream: everything the calls in between enqueue, including the gaps the
 * host-sequenced GMRES leaves, is inside the measured interval (bench.py times its K steps with it). */
int nsb_timer_start(nsb_ctx *ctx);
int nsb_timer_stop(nsb_ctx *ctx, double *elapsed_ms);
/* Launch counter of this context's own kernels (for bench.py gpu_launches). */
int64_t nsb_launch_count(const nsb_ctx *ctx);
/* timers of the last step in ms: [0] assemble, [1] prec init, [2] solve, [3] forces */
int nsb_timers(const nsb_ctx *ctx, double out_ms[4]);
/* info: [0] n_u [1] n_p [2] n_cells [3] nnz A00 (canonical) [4] nnz A01 [5] nnz A10 [6] nnz S
 *       [7] n_q [8] device bytes allocated [9] Chebyshev degree on F in effect
 *       [10] fine-level S sweeps per preconditioner application [11] Schur solver mode
 *       [12] number of levels of the Schur hierarchy
 *       [13] entries of the slab storage of F_s incl. padding [14] sum of the slab windows (nodes)
 *       [15] number of slabs [16] entries of the slab storage of A01 incl. padding
 *       [17] sum of the pressure windows
 *       [18] Gram-Schmidt re-orthogonalisation passes taken so far (third read of the Krylov basis)
 *       [19] bit 0: halo / all-gather exchanges run over peer memory (else NCCL), bit 1: the fine level of the
 *            Schur hierarchy is distributed over the ranks */
int nsb_info(const nsb_ctx *ctx, int64_t out[20]);

/* Parameters of the inner F polynomial in effect for the last step: [0] degree k, [1] lmax/lmin ratio of the real
 * interval, [2] lambda_max(D^-1 F) estimate, [3] imaginary half-axis of the ellipse (0 for a symmetric F). */
int nsb_inner_params(const nsb_ctx *ctx, double out[4]);

/* Host-only check of the slab (windowed sliced-ELL) storage the solver kernels stream F_s from
 * (csrc/slab.cuh): builds the layout from a node-level CSR pattern and evaluates y = (F_s (x) I_dim) x
 * on the host in exactly the order the device kernels use.  No device is touched.
 * stats: [0] slabs [1] stored entries [2] entries incl. padding [3] largest window [4] sum of windows
 *        [5] 1000 x shared-memory wavefronts per half-warp gather (1000 = free of bank conflicts). */
int nsb_slab_host_check(int dim, int64_t n_rows, int64_t n_cols, const int64_t *rowptr, const uint32_t *colind,
                        const double *val, uint32_t window_cap, const double *x, double *y, int64_t stats[6]);

/* Same for A01 in the slabs of the node pattern: y = A01 xp (n_nodes*dim rows) evaluated on the host in
 * kernel order.  stats: [0] stored entries [1] entries incl. padding [2] largest pressure window. */
int nsb_gslab_host_check(int dim, int64_t n_nodes, int64_t n_node_cols, const int64_t *node_rowptr,
                         const uint32_t *node_colind, uint32_t window_cap, const int64_t *rowptr01,
                         const uint32_t *colind01, const double *val01, const double *xp, double *y, int64_t stats[3]);

/* Host-only: coefficients of the degree-k Chebyshev-Jacobi polynomial the inner F solve of
 * PreconditionASIMPLE::vmult (reference src/NavierStokes.cpp:978-981) is replaced by, for the ellipse with
 * real interval [lmax/ratio, lmax] and imaginary half-axis imag (imag = 0: the interval polynomial):
 *   z_1 = inv_theta Dinv b,   z_{i+1} = z_i + c1[i] (z_i - z_{i-1}) + c2[i] Dinv (b - F z_i),  1 <= i < k.
 * c1, c2: k doubles each (entry 0 unused).  No device is touched. */
int nsb_cheb_coeffs_host_check(int k, double lmax, double ratio, double imag, double *inv_theta, double *c1, double *c2);

/* Host-only: largest singular value of the skew part of a dense m x m matrix (row-major) and its right singular
 * vector y (m doubles) -- the small eigenproblem behind the estimate of the imaginary extent of D^-1 F. */
int nsb_skew_radius_host_check(int m, const double *H, double *sigma, double *y);

/* Host-only: the greedy aggregation of the Schur hierarchy (csrc/amg.cuh: coarsen) on a CSR matrix.
 * measure / theta as in nsb_set_schur_strength; owner: optional (n entries, non-decreasing) -- aggregates never mix
 * owners.  agg_out: n entries, fine row -> aggregate; *n_coarse_out: number of aggregates; coarse_nnz_out: entries
 * of the Galerkin pattern.  No device is touched. */
int nsb_amg_coarsen_host_check(int64_t n, const int64_t *rowptr, const uint32_t *colind, const double *val, double theta,
                               int max_agg, int measure, const int32_t *owner, uint32_t *agg_out, int64_t *n_coarse_out,
                               int64_t *coarse_nnz_out);

/* Host-only: the reference-cell contraction tables the assembly kernel works from (csrc/fe_tables.h), in the
 * local dof order of deal.II's FE_SimplexP(2) / FE_SimplexP(1) (vertices, then lines (0,1),(1,2),(2,0)[,(0,3),(1,3),(2,3)]):
 *   mhat[a*nn+b]                 = int phi_a phi_b
 *   khat[(d*dim+e)*nn*nn + a*nn+b] = int d_d(phi_a) d_e(phi_b)
 *   chat[(n*dim+d)*nn*nn + a*nn+b] = int phi_a phi_n d_d(phi_b)
 *   dhat[(a*nv+k)*dim + d]        = int d_d(phi_a) psi_k
 * over the reference simplex, evaluated with quadrature rule `rule` (NSB_QUAD_*).  nn = 6 / 10, nv = dim + 1.
 * No device is touched. */
int nsb_fe_tables_host_check(int dim, int rule, double *mhat, double *khat, double *chat, double *dhat);

/* pinned host memory for callers that want asynchronous copies */
void *nsb_alloc_pinned(int64_t bytes);
void nsb_free_pinned(void *p);

/* ---- multi-GPU: one process per GPU of one box ------------
 * Domain decomposition of the reference (partition_triangulation + Epetra row
 * maps, NavierStokes.cpp:19-23, 71-86): velocity rows are distributed with a
 * ghost halo, pressure vectors and S are replicated.  A local vector is laid
 * out [u owned | u ghost | p] (see nsh_localize in nsb_host.h for the arrays).
 * Call order: nsb_create, nsb_set_mesh (local cells), nsb_set_local_dofs,
 * nsb_set_halo, nsb_comm_init, nsb_set_node_pattern / nsb_set_pattern with the
 * local patterns (A01: owned velocity rows, A10: owned pressure rows, S: full),
 * then as on one GPU.  nsb_set_dirichlet takes local dof ids of owned nodes;
 * nsb_set_solution / nsb_get_solution move the local vector.  Collectives:
 * NCCL grouped send/recv for the halo, all-reduce for dot products, S values
 * and the force integrals, broadcast groups to replicate pressure rows. */
/* 128-byte NCCL unique id created on rank 0 and broadcast by the caller. */
int nsb_comm_unique_id(char id[128]);
int nsb_comm_init(nsb_ctx *ctx, int rank, int n_ranks, const char id[128]);
/* p_offsets: n_ranks+1 offsets of the owned pressure ranges in the distributed
 * numbering; cell_nodes: local node ids, cell_pverts: distributed pressure ids. */
int nsb_set_local_dofs(nsb_ctx *ctx, int rank, int n_ranks, uint32_t n_own_nodes, uint32_t n_ghost_nodes,
                       uint32_t n_p, const uint32_t *p_offsets, const uint32_t *cell_nodes,
                       const uint32_t *cell_pverts);
/* For neighbour k: send the owned local nodes send_idx[send_ptr[k]..send_ptr[k+1])
 * and receive into the ghost slots [recv_ptr[k], recv_ptr[k+1]). */
int nsb_set_halo(nsb_ctx *ctx, int n_neighbors, const int32_t *neighbors, const int64_t *send_ptr,
                 const uint32_t *send_idx, const int64_t *recv_ptr);
/* Collects the owned velocity dofs of all ranks into one vector in the distributed numbering
 * (node_offsets: n_ranks+1 owned-node offsets; out_host: dim * node_offsets[n_ranks] doubles, same
 * content on every rank).  The pressure part of nsb_get_solution is already replicated. */
int nsb_gather_velocity(nsb_ctx *ctx, const uint32_t *node_offsets, double *out_host);

#ifdef __cplusplus
}
#endif
#endif
