/* Host-side setup library (no CUDA): what NavierStokes::setup() of the
 * reference obtains from deal.II/METIS, as plain arrays.
 *
 * Replaces (reference src/NavierStokes.cpp):
 *   :11-23   GridIn::read_msh + partition_triangulation   -> nsh_problem_read / _generate / nsh_partition
 *   :35-41   FESystem(P2^dim, P1)                          -> fixed Taylor-Hood P2/P1 tables
 *   :65-86   distribute_dofs + component_wise + index sets -> "cell_dofs", "cell_nodes", "cell_pverts"
 *   :101-117 make_sparsity_pattern (block CSR)             -> "a00.*", "a01.*", "a10.*", "s.*", "nodes.*"
 *   :297-324 interpolate_boundary_values                   -> "bc.dofs", "bc.values"
 *   :870-877 faces with boundary id 4                      -> "ff.cell", "ff.normal", "ff.measure"
 * All functions return 0 on success and a negative code on failure
 * (nsh_last_error() holds the message); nothing throws across the boundary.
 */
#ifndef NSB_HOST_H
#define NSB_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nsh_problem nsh_problem;

enum { NSH_INLET_PARABOLIC = 0, NSH_INLET_UNIFORM = 1 };

const char *nsh_last_error(void);

/* name: "2d-cylinder" (C1/C2), "3d-square" (C3), "3d-cylinder" (C5),
 * "naca2412" (C4), "channel2d", "channel3d"; h = target edge length. */
int nsh_problem_generate(const char *name, double h, nsh_problem **out);
/* gmsh ASCII .msh 2.2 / 4.1 (reference :11-17). */
/* The reference's airfoil pre-processing without gmsh (mesh/test.py:25-41, 155-168 + tests/2D/test_naca/run_test.sh:7-9:
 * `./test.py naca.dat 0.4 <angle>; gmsh NACA_2408.geo -2`): a contour from a .dat file in the mesh/naca.dat layout
 * (dat_path non-empty) or the analytic NACA 4-digit family (naca4, e.g. 2408), scaled to `chord`, turned clockwise
 * by aoa_deg about mid-chord, placed at (cx, cy) in the box [0,Lx] x [0,Ly] (test.py: 2.2 x 1.0, centre (0.4, 0.5));
 * boundary ids 0 bottom, 1 outlet, 2 top, 3 inlet, 4 airfoil. */
int nsh_problem_generate_airfoil(const char *dat_path, int naca4, double chord, double aoa_deg, double Lx, double Ly,
                                 double cx, double cy, double h, nsh_problem **out);
int nsh_problem_read(const char *msh_path, int dim, nsh_problem **out);
int nsh_problem_from_arrays(int dim, int64_t n_verts, const double *xyz, int64_t n_cells,
                            const uint32_t *cells, int64_t n_bfaces, const uint32_t *bfaces,
                            const int32_t *bids, nsh_problem **out);
int nsh_problem_write_msh(const nsh_problem *, const char *path);
void nsh_problem_free(nsh_problem *);

/* DoF numbering + patterns (reference :65-117).  expand_a00 = 0 skips the
 * canonical A00 expansion (the device can expand from "nodes.*"). */
int nsh_build_space(nsh_problem *, int expand_a00);
/* Inlet profile (drivers' InletVelocity), then the Dirichlet dof list and the
 * obstacle faces. */
int nsh_set_inlet(nsh_problem *, int kind, double U_m, double H, int time_sin);
int nsh_build_boundary(nsh_problem *);
/* InletVelocity::get_mean_vel() and the time factor of the inlet (1 or
 * sin(pi t/8)). */
double nsh_mean_velocity(const nsh_problem *, double time);
double nsh_inlet_time_factor(const nsh_problem *, double time);

/* sizes: [0] dim [1] n_verts [2] n_cells [3] n_bfaces [4] n_nodes [5] n_u
 * [6] n_p [7] dofs_per_cell [8] n_bc [9] n_force_faces */
int nsh_sizes(const nsh_problem *, int64_t out[10]);
/* Borrowed pointer to a named array (valid until the problem is freed):
 * "xyz" f64, "cells" u32, "bfaces" u32, "bids" i32, "cell_dofs" u32,
 * "cell_nodes" u32, "cell_pverts" u32, "node_xyz" f64,
 * "<blk>.rowptr" i64 / "<blk>.colind" u32 for blk in nodes,a00,a01,a10,s,
 * "bc.dofs" u32, "bc.values" f64, "ff.cell" u32, "ff.normal" f64,
 * "ff.measure" f64, "part.cell" i32 (after nsh_partition). */
int nsh_array(const nsh_problem *, const char *name, const void **data, int64_t *count, int *elem_bytes);

/* Recursive coordinate bisection of the cells into n_parts (stands in for
 * GridTools::partition_triangulation -> METIS, reference :19). */
int nsh_partition(nsh_problem *, int n_parts);

/* ---- domain decomposition over the GPUs of one box --------
 * Local view of rank `rank` out of n_parts after nsh_partition(n_parts):
 * owned + ghost velocity nodes in local numbering, replicated pressure in the
 * partition-contiguous ("distributed") numbering, local CSR patterns, halo
 * send/receive lists, owned Dirichlet nodes and obstacle faces.  Stands in for
 * parallel::fullydistributed::Triangulation + the Epetra maps the reference
 * builds at NavierStokes.cpp:19-23, 71-86, 113-127. */
typedef struct nsh_local nsh_local;
int nsh_localize(const nsh_problem *, int n_parts, int rank, nsh_local **out);
void nsh_local_free(nsh_local *);
/* sizes: [0] n_own [1] n_ghost [2] n_p [3] n_p_own [4] p_offset [5] n_local_cells
 * [6] n_neighbors [7] n_bc_nodes [8] n_force_faces [9] n_nodes_global [10] node_offset */
int nsh_local_sizes(const nsh_local *, int64_t out[11]);
/* "node_offset" "p_offset" "node_perm" "p_perm" "ghost_dist" "cells"
 * "cell_verts" "cell_nodes" "cell_pverts" (u32); "<blk>.rowptr" (i64) /
 * "<blk>.colind" (u32) for blk in fs,a01,a10,s; "neighbors" (i32) "send_ptr"
 * "recv_ptr" (i64) "send_idx" (u32); "bc_nodes" (u32) "bc_values" (f64);
 * "ff.cell" (u32) "ff.normal" "ff.measure" (f64). */
int nsh_local_array(const nsh_local *, const char *name, const void **data, int64_t *count, int *elem_bytes);

#ifdef __cplusplus
}
#endif
#endif
