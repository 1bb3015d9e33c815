// Facade implementation: the reference's method sequence (reference
// src/NavierStokes.cpp) expressed through include/nsb.h + the host setup
// library.  Console and CSV output follow the reference byte for byte where it
// is part of the de-facto interface (forces_vs_time.csv).
#include "NavierStokes.hpp"

#include <fcntl.h>
#include <sys/stat.h>

#include <chrono>
#include <cstdlib>
#include <iomanip>

#include <unistd.h>

#include <cstdio>
#include <thread>

#include "../../include/nsb.h"
#include "distribute.hpp"
#include "problem.hpp"

// Preconditioner of solve_time_step.  The reference selects it by (un)commenting blocks of
// NavierStokes.cpp:352-373; here it is a compile-time switch: NSB_PREC_ASIMPLE (the block the reference leaves
// enabled, :355-361), NSB_PREC_AYOSIDA (:364-373) or NSB_PREC_IDENTITY (:352).
#ifndef NS_PRECONDITIONER
#define NS_PRECONDITIONER NSB_PREC_ASIMPLE
#endif

namespace {
unsigned int env_uint(const char *name, unsigned int dflt) {
  const char *v = std::getenv(name);
  return v ? (unsigned int)std::strtoul(v, nullptr, 10) : dflt;
}
}  // namespace

NavierStokes::NavierStokes(const std::string &mesh_file_name_, const unsigned int &degree_velocity_,
                           const unsigned int &degree_pressure_, const double &deltat_, const double &T_,
                           const unsigned int &step_)
    : mesh_file_name(mesh_file_name_),
      mpi_size(env_uint("WORLD_SIZE", 1)),
      mpi_rank(env_uint("RANK", 0)),
      degree_velocity(degree_velocity_),
      degree_pressure(degree_pressure_),
      deltat(deltat_),
      T(T_),
      step(step_) {
  if (degree_velocity != 2 || degree_pressure != 1)
    throw std::invalid_argument("NavierStokes: only the Taylor-Hood pair P2/P1 of the reference drivers is built");
}

NavierStokes::~NavierStokes() {
  if (ctx) nsb_destroy(ctx);
}

void NavierStokes::check(int rc, const char *what) const {
  if (rc == NSB_OK) return;
  // the reference lets deal.II exceptions escape to the driver
  throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " +
                           (ctx ? nsb_last_error(ctx) : "no context"));
}

// reference :4-131
void NavierStokes::setup() {
  const bool pcout = mpi_rank == 0;
  if (pcout) std::cout << "Importing the mesh from " << mesh_file_name << std::endl;
  problem = std::make_unique<nsb::Problem>();
  problem->mesh = nsb::read_msh(mesh_file_name, dim);  // throws like GridIn::read_msh
  if (pcout) {
    std::cout << "\tNumber of elements = " << problem->mesh.n_cells() << std::endl;
    std::cout << "---------------------------------------------------" << std::endl;
    std::cout << "Initializing the finite element spaces" << std::endl;
  }
  problem->build_space(/*expand_a00=*/false);
  const nsb::DofMap &d = problem->dofs;
  if (pcout) {
    std::cout << "\tVelocity degree:\t\t" << degree_velocity << std::endl;
    std::cout << "\tPressure degree:\t\t" << degree_pressure << std::endl;
    std::cout << "\tDoFs per cell:\t\t\t" << d.dofs_per_cell() << std::endl;
    std::cout << "\tQuadrature points per cell:\t" << (dim == 2 ? 7 : (quad_rule == 0 ? 10 : 14)) << std::endl;
    std::cout << "\tQuadrature points per face:\t" << (dim == 2 ? 3 : 7) << std::endl;
    std::cout << "---------------------------------------------------" << std::endl;
    std::cout << "Initializing the DoF handler" << std::endl;
    std::cout << "\tNumber of DoFs:" << std::endl;
    std::cout << "\t\tvelocity:\t\t" << d.n_u << std::endl;
    std::cout << "\t\tpressure:\t\t" << d.n_p << std::endl;
    std::cout << "\t\ttotal:\t\t\t" << d.n_u + d.n_p << std::endl;
    std::cout << "---------------------------------------------------" << std::endl;
    std::cout << "Initializing the linear system" << std::endl;
    std::cout << "\tInitializing the sparsity pattern" << std::endl;
  }
  // boundary lists: dof set once, values re-evaluated per step through the
  // driver's InletVelocity (reference :297-324)
  problem->bfaces = nsb::boundary_faces(problem->mesh);
  problem->ff = nsb::force_faces(problem->mesh, problem->bfaces, 4);
  refresh_dirichlet(0.0);

  const int device = (int)env_uint("LOCAL_RANK", 0);
  if (nsb_create(dim, device, &ctx) != NSB_OK)
    throw std::runtime_error("NavierStokes::setup: no usable CUDA device (there is no CPU path)");
  if (mpi_size > 1) {
    setup_distributed();
    if (pcout) {
      std::cout << "\tInitializing the system right-hand side" << std::endl;
      std::cout << "\tInitializing the solution vector" << std::endl;
    }
    solution.assign((size_t)d.n_u + d.n_p, 0.0);
    return;
  }
  const nsb::Mesh &m = problem->mesh;
  const nsb::Patterns &P = problem->pat;
  check(nsb_set_mesh(ctx, (int64_t)m.n_verts(), m.xyz.data(), (int64_t)m.n_cells(), m.cells.data()), "nsb_set_mesh");
  check(nsb_set_dofs(ctx, d.n_u, d.n_p, d.cell_dofs.data()), "nsb_set_dofs");
  if (pcout) std::cout << "\tInitializing the matrices" << std::endl;
  check(nsb_set_node_pattern(ctx, P.nodes.n_rows, P.nodes.rowptr.data(), P.nodes.colind.data()), "nsb_set_node_pattern");
  check(nsb_set_pattern(ctx, NSB_A01, P.a01.n_rows, P.a01.rowptr.data(), P.a01.colind.data()), "nsb_set_pattern(A01)");
  check(nsb_set_pattern(ctx, NSB_A10, P.a10.n_rows, P.a10.rowptr.data(), P.a10.colind.data()), "nsb_set_pattern(A10)");
  check(nsb_set_pattern(ctx, NSB_S, P.s.n_rows, P.s.rowptr.data(), P.s.colind.data()), "nsb_set_pattern(S)");
  check(nsb_set_quadrature(ctx, quad_rule), "nsb_set_quadrature");
  check(nsb_set_force_faces(ctx, (int64_t)problem->ff.cell.size(), problem->ff.cell.data(), problem->ff.normal.data(),
                            problem->ff.measure.data()),
        "nsb_set_force_faces");
  check(nsb_set_params(ctx, deltat, nu), "nsb_set_params");
  check(nsb_set_solver(ctx, opt_rtol, opt_restart, opt_max_it, 0.5, NS_PRECONDITIONER), "nsb_set_solver");
  if (opt_sweeps_F > 0 && opt_sweeps_S > 0)
    check(nsb_set_inner(ctx, opt_sweeps_F, 2.5 * opt_sweeps_F, opt_sweeps_S, 0.45 * opt_sweeps_S * opt_sweeps_S),
          "nsb_set_inner");
  check(nsb_finalize_setup(ctx), "nsb_finalize_setup");
  if (pcout) {
    std::cout << "\tInitializing the system right-hand side" << std::endl;
    std::cout << "\tInitializing the solution vector" << std::endl;
  }
  solution.assign((size_t)d.n_u + d.n_p, 0.0);
}

void NavierStokes::set_solver_options(double gmres_rtol, int restart, int max_it, int sweeps_F, int sweeps_S) {
  opt_rtol = gmres_rtol;
  opt_restart = restart;
  opt_max_it = max_it;
  opt_sweeps_F = sweeps_F;
  opt_sweeps_S = sweeps_S;
  if (ctx) {
    check(nsb_set_solver(ctx, opt_rtol, opt_restart, opt_max_it, 0.5, NS_PRECONDITIONER), "nsb_set_solver");
    if (sweeps_F > 0 && sweeps_S > 0)
      check(nsb_set_inner(ctx, sweeps_F, 2.5 * sweeps_F, sweeps_S, 0.45 * sweeps_S * sweeps_S), "nsb_set_inner");
  }
}

// reference :297-324: the dof set is fixed, the values follow the inlet's time
void NavierStokes::refresh_dirichlet(double t) {
  inlet_velocity.set_time(t);
  const InletVelocity &inlet = inlet_velocity;
  problem->bc = nsb::dirichlet_dofs(problem->mesh, problem->dofs, problem->bfaces, [&](const double *x, int c) {
    Point<dim> p;
    for (unsigned int r = 0; r < dim; ++r) p[r] = x[r];
    return inlet.value(p, (unsigned int)c);
  });
  if (!ctx) return;
  if (mpi_size == 1) {
    check(nsb_set_dirichlet(ctx, (int64_t)problem->bc.dofs.size(), problem->bc.dofs.data(), problem->bc.values.data()),
          "nsb_set_dirichlet");
    return;
  }
  if (!local) return;
  // owned Dirichlet nodes of this rank, in local dof ids (the node set is fixed; only the values move)
  const nsb::DofMap &d = problem->dofs;
  std::vector<uint32_t> dofs;
  std::vector<double> vals;
  const uint32_t off = local->node_offset[mpi_rank], end = local->node_offset[mpi_rank + 1];
  for (size_t i = 0; i + dim <= problem->bc.dofs.size(); i += dim) {
    const uint32_t g = local->node_perm[problem->bc.dofs[i] / dim];
    if (g < off || g >= end) continue;
    for (unsigned int c = 0; c < dim; ++c) {
      dofs.push_back(dim * (g - off) + c);
      vals.push_back(problem->bc.values[i + c]);
    }
  }
  (void)d;
  check(nsb_set_dirichlet(ctx, (int64_t)dofs.size(), dofs.data(), vals.data()), "nsb_set_dirichlet");
}

// One process per GPU: RANK / WORLD_SIZE / LOCAL_RANK from the launcher (torchrun-style).  The NCCL id is
// handed from rank 0 to the others through a file named after the common parent process.
void NavierStokes::setup_distributed() {
  problem->has_boundary = true;  // bfaces / ff / bc were built in setup()
  problem->partition((int)mpi_size);
  local = std::make_unique<nsb::LocalProblem>(nsb::localize(*problem, (int)mpi_size, (int)mpi_rank));
  const nsb::LocalProblem &L = *local;
  const nsb::Mesh &m = problem->mesh;
  check(nsb_set_mesh(ctx, (int64_t)m.n_verts(), m.xyz.data(), (int64_t)L.cells.size(), L.cell_verts.data()), "nsb_set_mesh");
  check(nsb_set_local_dofs(ctx, (int)mpi_rank, (int)mpi_size, L.n_own, L.n_ghost, L.n_p, L.p_offset.data(),
                           L.cell_nodes.data(), L.cell_pverts.data()),
        "nsb_set_local_dofs");
  check(nsb_set_halo(ctx, (int)L.neighbors.size(), L.neighbors.data(), L.send_ptr.data(), L.send_idx.data(),
                     L.recv_ptr.data()),
        "nsb_set_halo");
  // The 128-byte NCCL id goes from rank 0 to the others through a file.  Its name carries a per-launch nonce
  // (the launcher's run id / rendezvous port, else the parent pid) so that a stale file of a crashed run is never
  // read; rank 0 creates it exclusively (0600) under a temporary name and renames it into place.
  char id[128];
  const char *rdv = std::getenv("NSB_RENDEZVOUS");
  std::string nonce;
  for (const char *k : {"TORCHELASTIC_RUN_ID", "MASTER_PORT"})
    if (const char *v = std::getenv(k)) nonce += std::string("_") + v;
  const std::string path = rdv ? rdv : "/tmp/nsb_nccl_id_" + std::to_string((long)getuid()) + nonce + "_" + std::to_string((long)getppid());
  if (mpi_rank == 0) {
    if (nsb_comm_unique_id(id) != NSB_OK) throw std::runtime_error("nsb_comm_unique_id failed (libnccl.so.2?)");
    const std::string tmp = path + ".tmp";
    ::unlink(tmp.c_str());
    ::unlink(path.c_str());  // left behind by a run that died between rename and remove
    const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL, 0600);
    if (fd < 0) throw std::runtime_error("cannot create " + tmp);
    const bool ok = ::write(fd, id, 128) == 128;
    ::close(fd);
    if (!ok || std::rename(tmp.c_str(), path.c_str()) != 0) {
      ::unlink(tmp.c_str());
      throw std::runtime_error("cannot write the NCCL id to " + path);
    }
  } else {
    bool got = false;
    // only a file younger than this process can belong to this launch
    const auto started = std::chrono::system_clock::now() - std::chrono::seconds(600);
    for (int tries = 0; tries < 2400 && !got; ++tries) {
      struct stat sb;
      if (::stat(path.c_str(), &sb) == 0 && sb.st_size == 128 &&
          std::chrono::system_clock::from_time_t(sb.st_mtime) > started) {
        if (FILE *f = std::fopen(path.c_str(), "rb")) {
          got = std::fread(id, 1, 128, f) == 128;
          std::fclose(f);
        }
      }
      if (!got) std::this_thread::sleep_for(std::chrono::milliseconds(50));
    }
    if (!got) throw std::runtime_error("timed out (120 s) waiting for the NCCL id of rank 0 at " + path);
  }
  check(nsb_comm_init(ctx, (int)mpi_rank, (int)mpi_size, id), "nsb_comm_init");
  if (mpi_rank == 0) std::remove(path.c_str());  // every rank has joined once comm_init returns
  check(nsb_set_node_pattern(ctx, L.fs.n_rows, L.fs.rowptr.data(), L.fs.colind.data()), "nsb_set_node_pattern");
  check(nsb_set_pattern(ctx, NSB_A01, L.a01.n_rows, L.a01.rowptr.data(), L.a01.colind.data()), "nsb_set_pattern(A01)");
  check(nsb_set_pattern(ctx, NSB_A10, L.a10.n_rows, L.a10.rowptr.data(), L.a10.colind.data()), "nsb_set_pattern(A10)");
  check(nsb_set_pattern(ctx, NSB_S, L.s.n_rows, L.s.rowptr.data(), L.s.colind.data()), "nsb_set_pattern(S)");
  check(nsb_set_quadrature(ctx, quad_rule), "nsb_set_quadrature");
  check(nsb_set_force_faces(ctx, (int64_t)L.ff_cell.size(), L.ff_cell.data(), L.ff_normal.data(), L.ff_measure.data()),
        "nsb_set_force_faces");
  check(nsb_set_params(ctx, deltat, nu), "nsb_set_params");
  check(nsb_set_solver(ctx, opt_rtol, opt_restart, opt_max_it, 0.5, NS_PRECONDITIONER), "nsb_set_solver");
  refresh_dirichlet(0.0);
  check(nsb_finalize_setup(ctx), "nsb_finalize_setup");
  local_vec.assign((size_t)dim * (L.n_own + L.n_ghost) + L.n_p, 0.0);
}

void NavierStokes::push_solution() {
  if (mpi_size == 1) {
    check(nsb_set_solution(ctx, solution.data()), "nsb_set_solution");
    return;
  }
  const nsb::LocalProblem &L = *local;
  const nsb::DofMap &d = problem->dofs;
  std::vector<uint32_t> canon(d.n_nodes);  // distributed id -> canonical node
  for (uint32_t A = 0; A < d.n_nodes; ++A) canon[L.node_perm[A]] = A;
  const uint32_t off = L.node_offset[mpi_rank];
  for (uint32_t i = 0; i < L.n_own + L.n_ghost; ++i) {
    const uint32_t A = canon[i < L.n_own ? off + i : L.ghost_dist[i - L.n_own]];
    for (unsigned int c = 0; c < dim; ++c) local_vec[(size_t)dim * i + c] = solution[(size_t)dim * A + c];
  }
  const size_t pu = (size_t)dim * (L.n_own + L.n_ghost);
  for (uint32_t V = 0; V < d.n_p; ++V) local_vec[pu + L.p_perm[V]] = solution[(size_t)d.n_u + V];
  check(nsb_set_solution(ctx, local_vec.data()), "nsb_set_solution");
}

void NavierStokes::pull_solution() {
  if (mpi_size == 1) {
    check(nsb_get_solution(ctx, solution.data()), "nsb_get_solution");
    return;
  }
  const nsb::LocalProblem &L = *local;
  const nsb::DofMap &d = problem->dofs;
  std::vector<double> g((size_t)d.n_u);
  check(nsb_gather_velocity(ctx, L.node_offset.data(), g.data()), "nsb_gather_velocity");
  check(nsb_get_solution(ctx, local_vec.data()), "nsb_get_solution");
  for (uint32_t A = 0; A < d.n_nodes; ++A)
    for (unsigned int c = 0; c < dim; ++c) solution[(size_t)dim * A + c] = g[(size_t)dim * L.node_perm[A] + c];
  const size_t pu = (size_t)dim * (L.n_own + L.n_ghost);
  for (uint32_t V = 0; V < d.n_p; ++V) solution[(size_t)d.n_u + V] = local_vec[pu + L.p_perm[V]];
}

// reference :133-330
void NavierStokes::assemble(const double &t) {
  forcing_term.set_time(t);
  refresh_dirichlet(t);
  check(nsb_set_params(ctx, deltat, nu), "nsb_set_params");
  check(nsb_assemble(ctx, t), "nsb_assemble");
}

// reference :332-341
void NavierStokes::set_re_number(int Re) {
  const bool pcout = mpi_rank == 0;
  if (pcout) std::cout << "-----------------------------------" << std::endl;
  const double U = inlet_velocity.get_mean_vel();
  nu = (U * Diameter) / Re;
  if (pcout) {
    std::cout << "New reynolds number setted to " << Re << " with nu = " << nu << " ." << std::endl;
    std::cout << "-----------------------------------" << std::endl;
  }
}

// reference :344-397
void NavierStokes::solve_time_step(std::ostream &oss) {
  int iters = 0;
  double t_prec = 0, t_sol = 0;
  check(nsb_solve_time_step(ctx, &iters, &t_prec, &t_sol), "nsb_solve_time_step");  // NSB_ENOCONV -> throws
  last_iters = (unsigned int)iters;
  if (mpi_rank == 0) {
    std::cout << "  " << iters << " GMRES iterations" << std::endl;
    std::cout << "Elapsed time for preconditioner initialisation: " << t_prec << " [s]" << std::endl;
    std::cout << "Elapsed time for time step solution: " << t_sol << " [s]" << std::endl;
    std::cout << std::endl;
  }
  oss << iters << "," << t_prec << "," << t_sol << ",";
  pull_solution();  // solution = solution_owned (:395)
}

// reference :831-929
void NavierStokes::compute_forces(const double & /*time*/) {
  if (mpi_rank == 0) std::cout << "Computing forces: " << std::endl;
  double out[4];
  check(nsb_compute_forces(ctx, inlet_velocity.get_mean_vel(), out), "nsb_compute_forces");
  drag = out[0];
  lift = out[1];
  cd = out[2];
  cl = out[3];
  if (mpi_rank == 0) {
    std::cout << "Drag coefficient (Cd): " << cd << "   Lift coefficient (Cl): " << cl << std::endl;
    std::cout << "---------------------------------------------------" << std::endl;
  }
}

// reference :439-499
void NavierStokes::solve(unsigned int time_step) {
  const bool pcout = mpi_rank == 0;
  if (pcout) std::cout << "===================================================" << std::endl;
  // The reference lets EVERY rank write ./forces_vs_time.csv (:446); the wall-clock columns differ between ranks,
  // so under mpirun the ranks overwrite each other's lines with text of different length.  Here rank 0 writes
  // the file and the other ranks write the same lines to a null stream.
  std::ofstream output_file(mpi_rank == 0 ? "forces_vs_time.csv" : "/dev/null");
  output_file << "time,deltat,GMRES_iters,time_prec_init,time_sol,Drag,Lift,Cd,Cl\n";
  if (0 == time_step) {
    time = 0.0;
    if (pcout) std::cout << "Applying initial conditions" << std::endl;
    std::fill(solution.begin(), solution.end(), 0.0);  // InitialConditions == 0
  } else {
    time = deltat * time_step;
    if (pcout) std::cout << "Continuing execution from time step " << time_step << std::endl;
    import_data(time_step);
  }
  push_solution();
  export_data(time_step);
  if (pcout) std::cout << "---------------------------------------------------" << std::endl;
  while (time < T - 0.5 * deltat) {
    time += deltat;
    ++time_step;
    if (pcout) std::cout << "n = " << std::setw(3) << time_step << ", t = " << std::setw(5) << time << ":" << std::flush;
    assemble(time);
    output_file << time << "," << deltat << ",";
    solve_time_step(output_file);
    compute_forces(time);
    output_file << drag << "," << lift << "," << cd << "," << cl << "\n";
    if (0 == time_step % step) {
      output(time_step);
      export_data(time_step);
    }
  }
  output_file.close();
}

// reference :571-784.  With one process the rank-count-independent order is the
// first-encounter order of the dofs when walking the cells by coarse-cell id
// and each cell's dof_indices in FESystem order.
void NavierStokes::compute_ordered_dofs_indices() {
  const nsb::DofMap &d = problem->dofs;
  const size_t N = (size_t)d.n_u + d.n_p;
  renumbered_dofs.assign(N, 0);
  std::vector<char> seen(N, 0);
  unsigned int k = 0;
  for (uint32_t dof : d.cell_dofs)
    if (!seen[dof]) {
      seen[dof] = 1;
      renumbered_dofs[dof] = k++;
    }
}

// reference :501-568: N raw doubles, position renumbered_dofs[i] holds dof i
void NavierStokes::export_data(const unsigned int &time_step) {
  if (renumbered_dofs.size() != solution.size()) compute_ordered_dofs_indices();
  if (mpi_rank != 0) return;
  std::vector<double> rbuf(solution.size());
  for (size_t i = 0; i < solution.size(); ++i) rbuf[renumbered_dofs[i]] = solution[i];
  const std::string file_name("../cache/state-ns-" + std::to_string(time_step) + ".dat");
  std::ofstream f(file_name, std::fstream::binary);
  f.write(reinterpret_cast<const char *>(rbuf.data()), (std::streamsize)(rbuf.size() * sizeof(double)));
  f.close();
  // a checkpoint that cannot be written must not be lost silently: a later solve(time_step) depends on it
  if (!f) throw std::runtime_error("export_data: cannot write " + file_name);
}

// reference :787-805
void NavierStokes::import_data(const unsigned int &time_step) {
  if (renumbered_dofs.size() != solution.size()) compute_ordered_dofs_indices();
  const std::string file_name("../cache/state-ns-" + std::to_string(time_step) + ".dat");
  std::ifstream f(file_name, std::fstream::binary);
  std::vector<double> inbuff(solution.size());
  f.read(reinterpret_cast<char *>(inbuff.data()), (std::streamsize)(inbuff.size() * sizeof(double)));
  if (!f) throw std::runtime_error("import_data: cannot read " + file_name);
  for (size_t i = 0; i < solution.size(); ++i) solution[i] = inbuff[renumbered_dofs[i]];
  if (ctx) push_solution();
}

// reference :808-828
void NavierStokes::post_process(const unsigned int &initial_time_step, const unsigned int &final_time_step,
                                const unsigned int &step_) {
  const bool pcout = mpi_rank == 0;
  if (pcout) std::cout << "=======================================================" << std::endl;
  for (unsigned int s = initial_time_step; s <= final_time_step; s += step_) {
    if (pcout) std::cout << "Importing time step " << s << " for post processing" << std::endl;
    import_data(s);
    compute_forces(s);
    if (pcout) std::cout << "Exporting pvtu files for time step " << s << std::endl;
    output(s);
  }
}

// reference :400-436.  DataOut writes one patch per cell (vertices are not
// shared between patches) with velocity / pressure / partitioning as point
// data; the same layout is written here as ASCII VTU plus the .pvtu record.
void NavierStokes::output(const unsigned int &time_step) const {
  const nsb::Mesh &m = problem->mesh;
  const nsb::DofMap &d = problem->dofs;
  const int nv = dim + 1, NN = d.nn();
  // each rank writes the cells its partition owns (one piece per rank, like write_vtu_with_pvtu_record)
  std::vector<size_t> mine;
  for (size_t c = 0; c < m.n_cells(); ++c)
    if (mpi_size == 1 || problem->part_cell[c] == (int)mpi_rank) mine.push_back(c);
  const size_t nc = mine.size();
  const std::string base = "output-stokes_" + std::to_string(time_step);
  const std::string piece = base + "." + std::to_string(mpi_rank) + ".vtu";
  std::ofstream f("../output/" + piece);
  if (!f) return;  // the reference's DataOut would throw; a missing ../output is not fatal for the solver
  f << std::setprecision(9);
  f << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
    << "<UnstructuredGrid>\n<Piece NumberOfPoints=\"" << nc * nv << "\" NumberOfCells=\"" << nc << "\">\n";
  f << "<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
  for (size_t ci = 0; ci < nc; ++ci)
    for (int a = 0; a < nv; ++a) {
      const size_t c = mine[ci];
      const double *p = &m.xyz[(size_t)m.cells[c * nv + a] * dim];
      f << p[0] << " " << p[1] << " " << (dim == 3 ? p[2] : 0.0) << "\n";
    }
  f << "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">\n";
  for (size_t i = 0; i < nc * nv; ++i) f << i << ((i + 1) % nv ? " " : "\n");
  f << "</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n";
  for (size_t c = 1; c <= nc; ++c) f << c * nv << "\n";
  f << "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n";
  for (size_t c = 0; c < nc; ++c) f << (dim == 2 ? 5 : 10) << "\n";
  f << "</DataArray>\n</Cells>\n<PointData Scalars=\"scalars\">\n";
  f << "<DataArray type=\"Float64\" Name=\"velocity\" NumberOfComponents=\"3\" format=\"ascii\">\n";
  for (size_t ci = 0; ci < nc; ++ci)
    for (int a = 0; a < nv; ++a) {
      const size_t node = d.cell_nodes[mine[ci] * NN + a];
      for (int k = 0; k < 3; ++k) f << (k < (int)dim ? solution[dim * node + k] : 0.0) << (k == 2 ? "\n" : " ");
    }
  f << "</DataArray>\n<DataArray type=\"Float64\" Name=\"pressure\" format=\"ascii\">\n";
  for (size_t ci = 0; ci < nc; ++ci)
    for (int a = 0; a < nv; ++a) f << solution[(size_t)d.n_u + d.cell_pverts[mine[ci] * nv + a]] << "\n";
  f << "</DataArray>\n<DataArray type=\"Float64\" Name=\"partitioning\" format=\"ascii\">\n";
  for (size_t ci = 0; ci < nc; ++ci)
    for (int a = 0; a < nv; ++a) f << (problem->part_cell.empty() ? 0 : problem->part_cell[mine[ci]]) << "\n";
  f << "</DataArray>\n</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
  if (mpi_rank == 0) {
    std::ofstream pv("../output/" + base + ".pvtu");
    pv << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
       << "<PUnstructuredGrid GhostLevel=\"0\">\n<PPointData Scalars=\"scalars\">\n"
       << "<PDataArray type=\"Float64\" Name=\"velocity\" NumberOfComponents=\"3\" format=\"ascii\"/>\n"
       << "<PDataArray type=\"Float64\" Name=\"pressure\" format=\"ascii\"/>\n"
       << "<PDataArray type=\"Float64\" Name=\"partitioning\" format=\"ascii\"/>\n</PPointData>\n"
       << "<PPoints>\n<PDataArray type=\"Float64\" NumberOfComponents=\"3\"/>\n</PPoints>\n";
    for (unsigned int r = 0; r < mpi_size; ++r) pv << "<Piece Source=\"" << base << "." << r << ".vtu\"/>\n";
    pv << "</PUnstructuredGrid>\n</VTKFile>\n";
  }
}
