// Domain decomposition of one Problem over the GPUs of a box: what the
// reference obtains from GridTools::partition_triangulation +
// parallel::fullydistributed::Triangulation + Epetra row maps (reference
// src/NavierStokes.cpp:19-23, 71-86, 113-127).
//
// * cells are partitioned (Problem::partition); a velocity node / pressure
//   vertex belongs to the lowest part touching it (deal.II rule);
// * "distributed numbering": each part owns a contiguous range of nodes and of
//   pressure vertices, canonical relative order inside a part (this is deal.II's
//   P-rank numbering);
// * a rank assembles every cell that touches one of its owned nodes (its owned
//   cells plus one layer), so owned matrix rows are complete without any
//   compress(add) exchange;
// * velocity vectors are row-distributed with a ghost halo; pressure vectors and
//   the Schur matrix S are replicated (n_p is ~4 % of the unknowns).
#pragma once
#include <cstdint>
#include <vector>

#include "problem.hpp"

namespace nsb {

struct LocalProblem {
  int rank = 0, n_parts = 1, dim = 0;
  uint32_t n_nodes_global = 0, n_p = 0;
  std::vector<uint32_t> node_offset, p_offset;  // n_parts+1: owned ranges in the distributed numbering
  std::vector<uint32_t> node_perm, p_perm;      // canonical id -> distributed id (global, same on all ranks)
  uint32_t n_own = 0, n_ghost = 0;              // local velocity nodes: [0,n_own) owned, then ghosts
  std::vector<uint32_t> ghost_dist;             // distributed ids of the ghosts, ascending (grouped by owner)
  std::vector<uint32_t> cells;                  // global ids of the local cells, ascending
  std::vector<uint32_t> cell_verts;             // n_loc_cells*(dim+1), global vertex ids
  std::vector<uint32_t> cell_nodes;             // n_loc_cells*NN, local node ids
  std::vector<uint32_t> cell_pverts;            // n_loc_cells*(dim+1), distributed pressure ids
  Csr fs;                                       // n_own x (n_own+n_ghost), node level
  Csr a01;                                      // dim*n_own x n_p (distributed pressure ids)
  Csr a10;                                      // n_p_own x dim*(n_own+n_ghost)
  Csr s;                                        // n_p x n_p, distributed numbering, replicated
  // halo: for neighbour k, send the owned local nodes send_idx[send_ptr[k]..send_ptr[k+1]) and receive the
  // ghost slots n_own + [recv_ptr[k], recv_ptr[k+1])
  std::vector<int32_t> neighbors;
  std::vector<int64_t> send_ptr, recv_ptr;
  std::vector<uint32_t> send_idx;
  // boundary data restricted to this rank
  std::vector<uint32_t> bc_nodes;  // owned local node ids, ascending
  std::vector<double> bc_values;   // dim per node (profile values, time factor 1)
  std::vector<uint32_t> ff_cell;   // local cell index of the obstacle faces in cells this rank owns
  std::vector<double> ff_normal, ff_measure;
  uint32_t n_p_own() const { return p_offset[rank + 1] - p_offset[rank]; }
};

// `p` must have its space and boundary lists built and be partitioned into
// n_parts (p.part_cell).
LocalProblem localize(const Problem &p, int n_parts, int rank);

}  // namespace nsb
