// Host-side structure analysis of the pose-graph solver (no CUDA in this header).
//
// g2o's BlockSolver::buildStructure + LinearSolverCSparse's symbolic step (SURVEY.md appendix C5,
// C6, C8) decide, once per graph structure, where every 3x3 block of H lives and in which order
// the unknowns are eliminated. This file does the same job for the GPU factorisation:
//   * fill-reducing ordering of the free vertices (nested dissection on the block graph; g2o
//     uses scalar AMD, SURVEY C8 -- the ordering changes rounding only, not the solution);
//   * elimination tree, block structure of the factor, column levels;
//   * an explicit, ordered list of every block update  M(i,j) -= M(i,k) D(k)^-1 M(j,k)^T  grouped
//     by the phase in which its source column k becomes final ("inspector"), which the kernels in
//     pgo_kernels.cu execute without searching and without atomics ("executor").
#ifndef CGM_PGO_SYMBOLIC_H
#define CGM_PGO_SYMBOLIC_H

#include <cstdint>
#include <string>
#include <vector>

namespace pgo {

// One block update  M(target) -= M(a) * Dinv(col_of(a)) * M(b)^T.  Updates of one phase are sorted
// by target; a thread owns a maximal run of equal targets (runs are short: the sources of one
// phase rarely share a target). kFinalFlag marks runs that complete the diagonal block of a
// column whose level equals the phase: the owner then also inverts it.
struct UpdateOp {
  int target;  // position of M(i,j) | kFinalFlag
  int a;       // position of M(i,k)
  int b;       // position of M(j,k)
};
static const int kFinalFlag = 0x40000000;
// Four consecutive target blocks of one column whose runs are congruent (same sources, source
// blocks in consecutive positions) are processed by one thread: the lead run carries kTileLead,
// the three member runs kTileMember (their heads skip). Positions therefore use 28 bits.
static const int kTileLead = 0x20000000;
static const int kTileMember = 0x10000000;
static const int kPosMask = 0x0FFFFFFF;
static const int kPanelWidth = 8;  // chain columns whose trailing updates are applied together

struct SolveOp {
  int row;  // target block row | kFinalFlag
  int pos;  // position of M(row, k)
};

struct Symbolic {
  int n = 0;                    // free (non-fixed) vertices = block rows of H
  std::vector<int> perm;        // perm[p] = hessian index eliminated p-th
  std::vector<int> iperm;       // iperm[hessian index] = p
  std::vector<int> parent;      // elimination tree over permuted columns (-1 = root)
  std::vector<int> level;       // level[p]: 0 for leaves, 1 + max(level of children) otherwise
  int n_levels = 0;
  // factor structure, block CSC, diagonal first, rows ascending
  std::vector<int> col_ptr;     // n + 1
  std::vector<int> row_idx;     // nnzb
  std::vector<int> col_of;      // nnzb: column of each position
  // the same structure by rows (strictly lower part), for the forward substitution
  std::vector<int> row_ptr;     // n + 1
  std::vector<int> row_pos;     // position of M(j,k) for each k < j in row j, k ascending
  // columns grouped by level
  std::vector<int> level_ptr;   // n_levels + 1
  std::vector<int> level_cols;  // n
  // update schedule: phase l (1 <= l < n_levels) applies ops [phase_ptr[l], phase_ptr[l+1])
  std::vector<int> phase_ptr;   // n_levels + 1 (phase 0 is empty)
  std::vector<UpdateOp> ops;
  // forward-substitution schedule, same idea: phase l applies  z(i) -= M(i,k) u(k)  for every
  // off-diagonal block of the columns k of level l - 1, sorted by target row i; kFinalFlag marks
  // the runs after which row i (of level l) is complete
  std::vector<int> fwd_ptr;     // n_levels + 1
  std::vector<SolveOp> fwd_ops;
  // multi-GPU domain decomposition (world > 1): owner[p] = rank that eliminates column p, or -1
  // for the shared top separators, which are ordered last: columns [first_shared, n) are shared.
  // A column's elimination-tree descendants all have the same owner; no edge connects two
  // different owners. explicit_final lists, per level, the shared columns that no shared-source
  // update completes (their children are all owned by ranks).
  int world = 1;
  std::vector<int> owner;          // n
  int first_shared = 0;            // == n when world == 1
  int local_levels = 0;            // 1 + max level of an owned column
  int shared_min_level = 0;        // min level of a shared column (n_levels if none)
  std::vector<int> xfinal_ptr;     // n_levels + 1
  std::vector<int> xfinal_cols;
  // statistics
  int64_t nnzb = 0, n_ops = 0;
  int max_run = 0;              // longest run of updates sharing one target within a phase
  double analyse_seconds = 0.0;
};

// Block adjacency of the free vertices: edges given as pairs of hessian indices (both >= 0).
// Returns false (with *err) on inconsistent input. `ordering`: 0 = nested dissection (default),
// 1 = natural (testing).
// `world` > 1 (a power of two) additionally cuts the top log2(world) dissection levels into shared
// separators and assigns every other column to a rank.
bool analyse(int n, const std::vector<std::pair<int, int> >& edges, int ordering, int world,
             Symbolic* out, std::string* err);

}  // namespace pgo
#endif
