// Host-side structure analysis of the pose-graph solver (no CUDA in this header).
//
// g2o's BlockSolver::buildStructure + LinearSolverCSparse's symbolic step (SURVEY.md appendix C5,
// C6, C8) decide, once per graph structure, where every 3x3 block of H lives and in which order
// the unknowns are eliminated. This file does the same job for the GPU factorisation:
//   * fill-reducing ordering of the free vertices (nested dissection on the block graph; g2o
//     uses scalar AMD, SURVEY C8 -- the ordering changes rounding only, not the solution);
//   * elimination tree, block structure of the factor;
//   * supernodes, panels, the scatter tables of the panels' outer products and the task lists per
//     level that the kernels in pgo_kernels.cu execute without searching ("inspector / executor");
//   * for world > 1, the cut of the top dissection levels into shared separators and the owner
//     of every other column.
#ifndef CGM_PGO_SYMBOLIC_H
#define CGM_PGO_SYMBOLIC_H

#include <cstdint>
#include <string>
#include <vector>

#if defined(__CUDACC__)
#define PGO_HOST_DEVICE __host__ __device__
#else
#define PGO_HOST_DEVICE
#endif

namespace pgo {

// ---- supernodal view (single-GPU factorisation and solves) -------------------------------------
// A SUPERNODE is a maximal chain of columns with nested structure (column q+1 is q's parent and
// struct(q+1) = struct(q) \ {q+1}); its blocks form a dense trapezoid in the block-CSC storage:
// block (i, t) of the chain's row list lives at col_ptr[c0 + t] + (i - t). A PANEL is a piece of
// at most kPanelWidth columns of a supernode: the unit of the right-looking factorisation
// (factor the panel, then subtract its outer product from every later column with atomics).
static const int kPanelWidth = 16;
static const int kSmallWidth = 4;    // panels / supernodes this narrow are handled by one warp
static const int kSubstWarpWidth = 4;   // substitutions: supernodes this narrow are warp tasks (16, i.e. every
                                        // single-panel supernode, was measured: the row gather of a 16-column
                                        // supernode is too long for one warp, 0.2 ms slower for one graph)
static const int kSubstSmallDoubles = 320;  // ... launched apart from the narrow ones above this need
static const int kSmallRows = 32;    // ... if the panel also has at most this many rows below it
static const int kFusedSmallDoubles = 704;  // 5.5 kB per warp: 8+ CTAs of 4 warps per SM
static const int kFactorSmallDoubles = 5400;  // 43 kB: a panel of <= 8 columns with a full row chunk
static const int kRowChunk = 64;     // rows below a panel per CTA task of the panel factorisation
// outer-product tile of one CTA task (tensor-core GEMM, pgo_kernels.cu): ti x tj blocks with
// ti a multiple of 8 and tj in {8, 16, 24, 32}; operands staged as [3 ti][ld] and [3 tj][ld]
// doubles with ld = sn_tile_ld(w), plus ti * tj scatter positions.
static const int kTileSmemDoubles = 11600;
PGO_HOST_DEVICE inline constexpr int sn_tile_ld(int w) {  // >= 3 w rounded up to 4, and = 4 (mod 16)
  return ((3 * w + 3) & ~3) + ((4 - (((3 * w + 3) & ~3) % 16)) + 16) % 16;
}
PGO_HOST_DEVICE inline constexpr int sn_tile_doubles(int w, int ti, int tj) {
  return 3 * (ti + tj) * sn_tile_ld(w) + (ti * tj + 1) / 2;
}
// shared memory (doubles) of one warp task: a fused small panel (sn_task_fused: Dg[w*w*9] Di[w*9]
// pair table, us[3w], xs[3w][3m+1]) and a narrow supernode of the substitutions
// (sn_task_backward_small / _forward_small: xs[3W] us[3W] Dg[W*W*9] Di[W*9] red[3 * 32])
PGO_HOST_DEVICE inline constexpr int sn_fused_doubles(int w, int m) {
  return w * w * 9 + w * 9 + (kSmallWidth * (kSmallWidth + 1) / 2 + 1) / 2 + 3 * w + 3 * w * (3 * m + 1);
}
PGO_HOST_DEVICE inline constexpr int sn_subst_doubles(int W) { return 6 * W + 9 * W * W + 9 * W + 96; }
static const int kQuickOrderingVertices = 4096;  // at most this many free vertices: cheaper dissection cuts
static const int kMaxSuperWidth = 1008;   // at most 63 panels per supernode
static const int kUpdateGroup = 4;        // panels per tile task of an ancestors' update

struct Task {
  int id;      // panel or supernode
  int r0, r1;  // row range of the below-row list (tile tasks: first row / first column)
  int aux;     // tile tasks: tile rows | tile columns << 8 | panels summed over << 16 |
               // panels between the last summed one and the task's panel << 22 | inside << 28
};

// Geometry of a panel / supernode in the block-CSC storage. Because the structure is nested,
// column t of the chain starts at  base + t * (w + m) - t (t - 1) / 2  and holds the chain rows
// t .. w-1 followed by the m rows below: no col_ptr look-ups on the device.
struct PanelDesc {
  int c0, w, m;   // first column, width, rows below the panel (later chain columns included)
  int base;       // col_ptr[c0]
  int meta;       // offset of the panel's rows in colbase / tbl_off
  int scratch;    // block offset in the diagonal scratch area, or -1
  int sn_off;     // offset of c0 inside its supernode
  int sn;         // supernode
};
struct SuperDesc {
  int c0, W, m;   // first column, width, rows below the supernode
  int base;       // col_ptr[c0]
  int pn_begin, pn_end;
  int pad0, pad1;
};

struct Supernodal {
  int n_super = 0, n_panels = 0, n_plevels = 0, n_slevels = 0;
  std::vector<int> sn_first;   // n_super + 1: first column of each supernode
  std::vector<int> sn_pn_ptr;  // n_super + 1: its panels (consecutive)
  std::vector<int> pn_first;   // n_panels + 1
  std::vector<int> pn_sn;      // n_panels
  std::vector<PanelDesc> pn;   // n_panels
  std::vector<SuperDesc> sn;   // n_super
  std::vector<int> pn_owner, sn_owner;  // rank that eliminates it, -1 = shared separator (world > 1)
  // scatter tables of the outer-product update. For below-row b of panel K (a block column r_b of
  // some later panel q) and below-row a >= b:  position of M(r_a, r_b) =
  // colbase[pn_meta[K] + b] + tbl[tbl_off[pn_meta[K] + b] + a].
  std::vector<int> pn_meta;    // n_panels + 1
  std::vector<int> colbase, tbl_off, tbl;
  // panels whose factorisation is split over several CTA tasks write their diagonal part to a
  // scratch area first (the other tasks still read the unfactored blocks): offset in blocks or -1
  std::vector<int> pn_scratch;
  int64_t scratch_blocks = 0;
  // factorisation tasks by panel level: fused (factor + update) warp tasks for small panels,
  // CTA tasks for the rest (fa: factor a row chunk; fb: one tile of the outer product)
  std::vector<int> ff_ptr, fa_ptr, fb_ptr;  // n_plevels + 1
  std::vector<Task> ff, fa, fb;
  // substitution tasks by supernode level: ss = whole small supernode (warp), sa = triangular part
  // of a wide supernode (CTA), sb = backward gather over the rows below a wide supernode's panel
  // (warp, chunks of 64 rows), sf = forward scatter over the same rows (chunks of 32; only for
  // right-hand sides solved after the factorisation: during an iteration the forward substitution
  // rides along with the factorisation)
  std::vector<int> ss_ptr, sa_ptr, sf_ptr, sb_ptr;  // n_slevels + 1
  std::vector<Task> ss, sa, sf, sb;
  // what the host needs to launch the phases: level pointers and per-level shared-memory needs
  // what one stage of an iteration launches: the tasks of the panels / supernodes of one owner
  // (kAllOwners: everything, the single-GPU case), by level, with per-level shared-memory needs
  struct Lists {
    int n_plevels = 0, n_slevels = 0;
    std::vector<int> ff_ptr, fa_ptr, fb_ptr, ss_ptr, sa_ptr, sf_ptr, sb_ptr;
    std::vector<Task> ff, fa, fb, ss, sa, sf, sb;
    std::vector<int> fa_smem, fb_smem;  // doubles of shared memory of the largest task per level
    // factor tasks: per level the ones needing more than kFactorSmallDoubles come first
    // (fa_large[l] of them, fa_smem[l]); the rest (narrow panels) are launched apart with
    // fa_smem_small[l], so that four of them fit an SM instead of two
    std::vector<int> fa_smem_small, fa_large;
    // fused warp tasks: per level the few tasks needing more than kFusedSmallDoubles of shared
    // memory come first (ff_large[l] of them, stride ff_smem[l]) and are launched apart from the
    // many small ones (stride ff_smem_small[l]), which then run at full occupancy
    std::vector<int> ff_smem, ff_smem_small, ff_large;
    std::vector<int> sa_smem;
    // warp substitution tasks (ss): per level the wider ones first (ss_large[l] of them, stride
    // ss_smem[l] doubles per warp), then the narrow ones (stride ss_smem_small[l])
    std::vector<int> ss_smem, ss_smem_small, ss_large;
  };
  static const int kAllOwners = -2;
  Lists lists(int owner) const;
  int64_t update_blocks = 0;   // target blocks touched by all outer products (atomic 3x3 adds)
  double flops = 0.0;          // of one numeric factorisation
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
  // multi-GPU domain decomposition (world > 1): owner[p] = rank that eliminates column p, or -1
  // for the shared top separators, which are ordered last: columns [first_shared, n) are shared.
  // A column's elimination-tree descendants all have the same owner; no edge connects two
  // different owners.
  int world = 1;
  std::vector<int> owner;          // n
  int first_shared = 0;            // == n when world == 1
  int local_levels = 0;            // 1 + max level of an owned column
  int shared_min_level = 0;        // min level of a shared column (n_levels if none)
  Supernodal sn;
  // statistics
  int64_t nnzb = 0;
  int64_t n_ops = 0;            // 3x3 block products of one factorisation (sum over columns of m (m + 1) / 2)
  double analyse_seconds = 0.0;
};

// Block adjacency of the free vertices: edges given as pairs of hessian indices (both >= 0).
// Returns false (with *err) on inconsistent input. `ordering`: 0 = nested dissection (default),
// 1 = natural (testing).
// The explicit block-update schedule (ops) is only built for world > 1; the single-GPU path uses
// the supernodal tables (sn).
// `world` > 1 (a power of two) additionally cuts the top log2(world) dissection levels into shared
// separators and assigns every other column to a rank.
bool analyse(int n, const std::vector<std::pair<int, int> >& edges, int ordering, int world,
             Symbolic* out, std::string* err);

}  // namespace pgo
#endif
