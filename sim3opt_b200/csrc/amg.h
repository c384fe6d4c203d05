// amg.h -- multilevel (aggregation) preconditioner for the pose-graph PCG: host-side hierarchy.
//
// The linear system of one LM trial, (H + lambda I) x = b (BlockSolver::solve + LinearSolverEigen
// [EXT g2o]), is solved by PCG.  Block-Jacobi alone needs O(graph diameter)
// iterations once lambda is small; this adds a coarse-space correction built on the gauge
// near-null space of a pose graph: moving every pose of an aggregate rigidly with its root,
// delta_i = Ad(S_i S_root^-1) xi  (left-multiplicative tangent, VertexSim3Expmap::oplusImpl);
// for the 4-DoF scale+translation and 1-DoF scale graphs of the stepwise pipeline the same gauge
// (a right-multiplied world similarity) gives (s_i/s_root) diag(1, R_i R_root^T) and s_i/s_root.
//
//   z = D^-1 r  +  P0 * V(P0^T r)          fine level: additive (no extra product with H)
//   V = V(1,1)-cycle with damped block-Jacobi smoothing on the Galerkin operators
//       A_{l+1} = P_l^T A_l P_l,  coarsest level inverted densely.
//
// The hierarchy (aggregates, coarse patterns, Galerkin contributor lists) depends on the block
// structure only and is built once on the host; the values are recomputed on the device per LM
// trial (amg.cu).  Every list is ordered, so the device sums are reproducible.
// Partitioned solve (one process per GPU): aggregates never cross a rank's vertex range, the fine
// transfer is local, and level 1 and below are replicated on every rank (amg.cu, AmgState).
#pragma once
#include <stdint.h>
#include <vector>

#include "internal.h"

namespace s3o {

// Transfer level l -> l+1 and the pattern of level l+1.
struct AmgHostLevel {
    int n_fine = 0;                       // vertices of level l
    int n = 0;                            // vertices of level l+1 (aggregates)
    std::vector<int32_t> agg;             // [n_fine] aggregate of each level-l vertex
    std::vector<int32_t> root;            // [n] level-l vertex whose frame the aggregate moves with
    std::vector<int32_t> vid;             // [n] original vertex id of that root
    std::vector<int32_t> mem_ptr, mem_idx; // [n+1], [n_fine] members of each aggregate, ascending
    // level l+1 operator: full BSR (both triangles), columns ascending
    std::vector<int32_t> rowptr, colidx, blk_row, dpos;
    // Galerkin contributor lists, one per upper block (I <= J) of level l+1
    int nub = 0;
    std::vector<int32_t> gal_ptr;         // [nub+1]
    std::vector<int32_t> gal_ent;         // (level-l block index << 2) | flag
                                          //   flag 0: P_i^T A P_j, 1: its transpose, 2: both (i != j inside one aggregate)
    std::vector<int32_t> gal_i, gal_j;    // row / column vertex (level l) of each entry's block
    std::vector<int32_t> gal_out;         // [nub] position of (I,J) in the level-(l+1) block array
    std::vector<int32_t> gal_mirror;      // [nub] position of (J,I), -1 on the diagonal
    std::vector<int32_t> gal_I, gal_J;    // [nub]
    std::vector<int32_t> gal_order;       // [nub] upper blocks by descending list length (warps get equal work)
};

// Builds the hierarchy below the BSR-upper structure S (level 0).  Stops when a level has at
// most `coarsest_max` vertices, when aggregation stalls, or after `max_levels` transfers.
// seg > 0 (partitioned solve): level-0 aggregates stay inside the index segments [r*seg, (r+1)*seg).
void amg_build_hierarchy(const HostStructure &S, int coarsest_max, int max_levels, std::vector<AmgHostLevel> &levels,
                         int seg = 0);

}  // namespace s3o
