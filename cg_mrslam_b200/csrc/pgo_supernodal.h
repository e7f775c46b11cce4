// Supernodal numeric factorisation and substitution of the permuted block Hessian (single-GPU
// path of the pose-graph solver). Replaces, for one GPU, the work of g2o's
// LinearSolverCSparse::solve (selected by the reference at
// src/slam/graph_slam.cpp:45-53): numeric Cholesky + two triangular solves.
//
// Storage and arithmetic are those of the level-by-level path (pgo_kernels.cu): block CSC of 3x3
// blocks, UNSCALED blocks  M(i,j) = A(i,j) - sum_k M(i,k) Dinv(k) M(j,k)^T,  D(j) = M(j,j),
// Dinv(j) = D(j)^-1; forward  z_i -= M(i,k) u_k, u_k = Dinv_k z_k; backward
// x_j = u_j - Dinv_j sum_{i>j} M(i,j)^T x_i.  What changes is the schedule: columns are grouped
// into supernodes / panels (pgo_symbolic.h), a panel is factorised as a small dense problem in
// shared memory and its outer product is subtracted from later columns with fp64 atomics
// (right-looking), so the number of grid-wide barriers is the number of PANEL levels (~100 for the
// 50k-vertex benchmark graph) instead of the elimination-tree height (~1000), and the index
// traffic is one table entry per target block instead of one 12-byte record per block product.
//
// Every task function is written against a "group" (rank / size / sync) and uses neither warp
// shuffles nor per-thread state across a sync, so the SAME source runs
//   * on the device with a CTA group (__syncthreads) or a warp group (__syncwarp), and
//   * on the host with a one-thread group (tests/hostsim/pgo_hostsim.cpp, TEST ONLY), which is
//     how the tables and the task logic are checked on a GPU-less box.
#ifndef CGM_PGO_SUPERNODAL_H
#define CGM_PGO_SUPERNODAL_H

#include <cmath>
#include <cstddef>

#include "pgo_symbolic.h"

#if defined(__CUDACC__)
#define PGO_HD __host__ __device__ __forceinline__
#else
#define PGO_HD inline
#endif

namespace pgo {

struct SNView {
  // structure (pgo_symbolic.h)
  const int* row_idx;
  const PanelDesc* pn;
  const SuperDesc* sn;
  const int* colbase;
  const int* tbl_off;
  const int* tbl;
  // numeric
  double* M;        // [nnzb][9]
  double* Y;        // [nnzb][9] Y(a,t) = M(a,t) Dinv_t for the rows below big panels (the scaled
                    // operand of the outer products, written by the panel factorisation)
  double* Dinv;     // [n][9]
  double* z;        // [n][3] right-hand side, overwritten by the forward substitution
  double* u;        // [n][3]
  double* x;        // [n][3] must be zero before the backward substitution
  double* scratch;  // [scratch_blocks][9]
  int* status;      // [0] set to 1 when a pivot block is not positive definite
  // batch of problem instances with this structure: element strides between instances
  long long s_M, s_Dinv, s_vec, s_scratch;
  int s_status;     // 4 for a batch of graphs, 0 when the "instances" are right-hand sides of one graph
};

// The view of instance b of a batch (blockIdx.y on the device).
PGO_HD SNView sn_at_instance(SNView V, int b) {
  V.M += b * V.s_M;
  V.Y += b * V.s_M;
  V.Dinv += b * V.s_Dinv;
  V.z += b * V.s_vec;
  V.u += b * V.s_vec;
  V.x += b * V.s_vec;
  V.scratch += b * V.s_scratch;
  V.status += b * V.s_status;
  return V;
}

// doubles of shared memory a group needs for any task of its kind
static const int kPairDoubles = (kPanelWidth * (kPanelWidth + 1) / 2 + 1) / 2;  // int table
static const int kFactorSmemDoubles = kPanelWidth * kPanelWidth * 9 + kPanelWidth * 9 + kPairDoubles +
                                      3 * kPanelWidth +
                                      3 * kPanelWidth * (3 * kRowChunk + 1);  // factor task
static const int kTriSmemDoubles = 6 * kMaxSuperWidth + kPanelWidth * kPanelWidth * 9 + kPanelWidth * 9 +
                                   3 * 256;  // a wide supernode's vectors (substitution)
PGO_HD constexpr int sn_max3(int a, int b, int c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }
static const int kCtaSmemDoubles = sn_max3(kFactorSmemDoubles, kTileSmemDoubles, kTriSmemDoubles);
static const int kSmallPairDoubles = (kSmallWidth * (kSmallWidth + 1) / 2 + 1) / 2;
static const int kWarpSmemDoubles = sn_fused_doubles(kSmallWidth, kSmallRows);  // fused
static const int kWarpSubstDoubles = sn_subst_doubles(kSubstWarpWidth);

struct SeqGroup {  // host: one thread plays every rank in turn
  PGO_HD int rank() const { return 0; }
  PGO_HD int size() const { return 1; }
  PGO_HD void sync() const {}
};

// L2-coherent load / accumulate: panels are written by other SMs in the previous phase
PGO_HD double sn_ld(const double* p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(p);
#else
  return *p;
#endif
}
PGO_HD void sn_add(double* p, double v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(p, v);
#else
  *p += v;
#endif
}
PGO_HD void sn_fail(int* status) {
#if defined(__CUDA_ARCH__)
  atomicExch(status, 1);
#else
  *status = 1;
#endif
}

// Asynchronous 8-byte global -> shared copy (cp.async) and the wait for all of a thread's copies;
// the host stand-in copies at once. dst must point into the task's shared memory.
PGO_HD void sn_copy_async(double* dst, const double* src) {
#if defined(__CUDA_ARCH__)
  const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(src) : "memory");
#else
  *dst = *src;
#endif
}
PGO_HD void sn_copy_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// first position of column t of a chain whose first column starts at `base` with `len` blocks
PGO_HD int sn_colpos(int base, int len, int t) { return base + t * len - t * (t - 1) / 2; }

// Memory-level parallelism: every rank issues eight independent loads before it consumes any of
// them (a plain copy loop stalls on each load's L2 latency in turn). loc(i, &src, &dst).
template <class G, class F>
PGO_HD void sn_gather(const G& g, int n, F loc) {
  const int step = g.size();
  for (int i0 = g.rank(); i0 < n; i0 += 8 * step) {
    double v[8];
    double* dp[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + j * step;
      dp[j] = nullptr;
      if (i < n) {
        const double* sp;
        loc(i, &sp, &dp[j]);
        v[j] = sn_ld(sp);
      }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 8; ++j)
      if (dp[j]) *dp[j] = v[j];
  }
}

PGO_HD void sn_ld9(const double* p, double* m) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 9; ++k) m[k] = sn_ld(p + k);
}

// Inverse of a symmetric positive definite 3x3 (lower triangle read) by cofactors: one division on
// the dependency chain instead of three square roots and six divisions. Positive definiteness by
// Sylvester's criterion; false if a leading minor is not positive.
PGO_HD bool sn_spd_inverse(const double* a, double* inv) {
  const double a00 = a[0], a10 = a[3], a20 = a[6], a11 = a[4], a21 = a[7], a22 = a[8];
  const double c00 = a11 * a22 - a21 * a21;
  const double c10 = a20 * a21 - a10 * a22;
  const double c20 = a10 * a21 - a20 * a11;
  const double c11 = a00 * a22 - a20 * a20;
  const double c21 = a10 * a20 - a00 * a21;
  const double c22 = a00 * a11 - a10 * a10;
  const double det = a00 * c00 + a10 * c10 + a20 * c20;
  if (!(a00 > 0.0) || !(c22 > 0.0) || !(det > 0.0)) return false;
  const double r = 1.0 / det;
  inv[0] = c00 * r;
  inv[1] = inv[3] = c10 * r;
  inv[2] = inv[6] = c20 * r;
  inv[4] = c11 * r;
  inv[5] = inv[7] = c21 * r;
  inv[8] = c22 * r;
  return true;
}

// ---- diagonal part of a panel: w x w blocks, Dg[(i * w + t) * 9 + k], i >= t loaded -------------
template <class G>
PGO_HD void sn_load_diag(const G& g, const SNView& V, int base, int len, int w, double* Dg) {
  // lower triangle incl. diagonal, pair p = i (i + 1) / 2 + t; the upper part is written later
  const int n_pairs = w * (w + 1) / 2;
  sn_gather(g, n_pairs * 9, [&](int idx, const double** sp, double** dp) {
    const int p = idx / 9, k = idx % 9;
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= p) ++i;
    const int t = p - i * (i + 1) / 2;
    *sp = V.M + 9 * static_cast<size_t>(sn_colpos(base, len, t) + (i - t)) + k;
    *dp = Dg + (i * w + t) * 9 + k;
  });
}

// In-place block LDL^T of the diagonal part. On return (after a sync):
//   Dg[i][t], i >= t : final M(c0+i, c0+t);  Dg[s][t], s < t : G(s,t) = Dinv_s M(t,s)^T;  Di[t].
// pairs: scratch table of w (w + 1) / 2 ints, (i << 8) | t per lower-triangular block, COLUMN by
// column, so that the blocks still alive at step s (t > s) are a suffix of the table.
template <class G>
PGO_HD void sn_factor_diag(const G& g, const SNView& V, int w, double* Dg, double* Di, int* pairs) {
  const int n_pairs = w * (w + 1) / 2;
  for (int p = g.rank(); p < n_pairs; p += g.size()) {
    int t = 0;
    while ((t + 1) * w - (t + 1) * t / 2 <= p) ++t;  // column t starts at t w - t (t - 1) / 2
    pairs[p] = ((t + p - (t * w - t * (t - 1) / 2)) << 8) | t;
  }
  for (int s = 0; s < w; ++s) {
    if (g.rank() == 0) {
      if (!sn_spd_inverse(Dg + (s * w + s) * 9, Di + 9 * s)) {
        sn_fail(V.status);
        for (int k = 0; k < 9; ++k) Di[9 * s + k] = 0.0;
      }
    }
    g.sync();
    // G(s,t) = Dinv_s M(t,s)^T for t > s, into the upper part
    const double* d = Di + 9 * s;
    const int rem = w - s - 1;
    for (int idx = g.rank(); idx < rem * 9; idx += g.size()) {
      const int t = s + 1 + idx / 9, k = idx % 9, r = k / 3, c = k % 3;
      const double* m = Dg + (t * w + s) * 9;
      Dg[(s * w + t) * 9 + k] = d[3 * r] * m[3 * c] + d[3 * r + 1] * m[3 * c + 1] + d[3 * r + 2] * m[3 * c + 2];
    }
    g.sync();
    // trailing update of the blocks (i, t), i >= t > s (a suffix of the pair table)
    const int first = (s + 1) * w - (s + 1) * s / 2;  // first pair of column s + 1
    for (int idx = 9 * first + g.rank(); idx < n_pairs * 9; idx += g.size()) {
      const int pr = pairs[idx / 9], i = pr >> 8, t = pr & 255;
      const int k = idx % 9, r = k / 3, c = k % 3;
      const double* a = Dg + (i * w + s) * 9;
      const double* gg = Dg + (s * w + t) * 9;
      Dg[(i * w + t) * 9 + k] -= a[3 * r] * gg[c] + a[3 * r + 1] * gg[3 + c] + a[3 * r + 2] * gg[6 + c];
    }
    g.sync();
  }
}

// Rows below the diagonal part, staged as scalar rows: xs[(3 t + c) * ldx + 3 a + r] = M(a,t)[r][c],
// as asynchronous copies, column by column (no division by a run-time value): the caller overlaps
// them with the factorisation of the diagonal part and calls sn_copy_wait() + sync before the rows
// are used.
template <class G>
PGO_HD void sn_load_rows_async(const G& g, const SNView& V, const PanelDesc& pd, int r0, int nrows, int ldx,
                               double* xs) {
  for (int t = 0; t < pd.w; ++t) {
    const double* src = V.M + 9 * static_cast<size_t>(sn_colpos(pd.base, pd.w + pd.m, t) + (pd.w - t) + r0);
    double* dst = xs + 3 * t * ldx;
    for (int q = g.rank(); q < 9 * nrows; q += g.size()) {
      const int a = q / 9, k = q - 9 * a, r = k / 3, c = k - 3 * r;
      sn_copy_async(dst + c * ldx + 3 * a + r, src + q);
    }
  }
}

template <class G>
PGO_HD void sn_store_rows(const G& g, const SNView& V, const PanelDesc& pd, int r0, int nrows, int ldx,
                          const double* xs) {
  for (int t = 1; t < pd.w; ++t) {  // column 0 is never modified inside the panel
    double* dst = V.M + 9 * static_cast<size_t>(sn_colpos(pd.base, pd.w + pd.m, t) + (pd.w - t) + r0);
    for (int idx = g.rank(); idx < 9 * nrows; idx += g.size()) {
      const int a = idx / 9, k = idx % 9;
      dst[idx] = xs[(3 * t + k % 3) * ldx + 3 * a + k / 3];
    }
  }
}

// Y(a,t) = M(a,t) Dinv_t for the staged rows: the scaled operand of the panel's outer products.
template <class G>
PGO_HD void sn_store_scaled_rows(const G& g, const SNView& V, const PanelDesc& pd, int r0, int nrows,
                                 int ldx, const double* xs, const double* Di) {
  for (int t = 0; t < pd.w; ++t) {
    double* dst = V.Y + 9 * static_cast<size_t>(sn_colpos(pd.base, pd.w + pd.m, t) + (pd.w - t) + r0);
    const double* d = Di + 9 * t;
    for (int a = g.rank(); a < nrows; a += g.size()) {
      for (int r = 0; r < 3; ++r) {
        const double m0 = xs[(3 * t) * ldx + 3 * a + r], m1 = xs[(3 * t + 1) * ldx + 3 * a + r],
                     m2 = xs[(3 * t + 2) * ldx + 3 * a + r];
        for (int c = 0; c < 3; ++c) dst[9 * a + 3 * r + c] = m0 * d[c] + m1 * d[3 + c] + m2 * d[6 + c];
      }
    }
  }
}

// M(a,t) -= sum_{s<t} M(a,s) G(s,t), one scalar row per rank.
template <class G>
PGO_HD void sn_solve_rows(const G& g, int w, const double* Dg, double* xs, int ldx, int n_scalar) {
  for (int row = g.rank(); row < n_scalar; row += g.size()) {
    for (int t = 1; t < w; ++t) {
      double a0 = xs[(3 * t) * ldx + row], a1 = xs[(3 * t + 1) * ldx + row],
             a2 = xs[(3 * t + 2) * ldx + row];
      for (int s = 0; s < t; ++s) {
        const double x0 = xs[(3 * s) * ldx + row], x1 = xs[(3 * s + 1) * ldx + row],
                     x2 = xs[(3 * s + 2) * ldx + row];
        const double* gg = Dg + (s * w + t) * 9;
        a0 -= x0 * gg[0] + x1 * gg[3] + x2 * gg[6];
        a1 -= x0 * gg[1] + x1 * gg[4] + x2 * gg[7];
        a2 -= x0 * gg[2] + x1 * gg[5] + x2 * gg[8];
      }
      xs[(3 * t) * ldx + row] = a0;
      xs[(3 * t + 1) * ldx + row] = a1;
      xs[(3 * t + 2) * ldx + row] = a2;
    }
  }
}

template <class G>
PGO_HD void sn_store_diag(const G& g, const SNView& V, const PanelDesc& pd, const double* Dg,
                          const double* Di, int scratch_at) {
  const int w = pd.w;
  if (scratch_at >= 0) {
    double* dst = V.scratch + 9 * static_cast<size_t>(scratch_at);
    for (int idx = g.rank(); idx < w * w * 9; idx += g.size()) dst[idx] = Dg[idx];
  } else {
    for (int idx = g.rank(); idx < w * w * 9; idx += g.size()) {
      const int i = idx / (w * 9), t = (idx / 9) % w, k = idx % 9;
      if (i >= t) V.M[9 * static_cast<size_t>(sn_colpos(pd.base, w + pd.m, t) + (i - t)) + k] = Dg[idx];
    }
  }
  for (int idx = g.rank(); idx < w * 9; idx += g.size())
    V.Dinv[9 * static_cast<size_t>(pd.c0) + idx] = Di[idx];
}

// Forward substitution rides along with the factorisation: the right-hand side is one more scalar
// row of the panel (z^T behaves like a block row of M under the unscaled update rule), staged at
// row index 3 * nrows of xs. After sn_solve_rows that row holds the final z of the panel's
// columns; u_t = Dinv_t z_t, and every row a below the panel gets  z_{r_a} -= sum_t M(a,t) u_t.
template <class G>
PGO_HD void sn_load_rhs(const G& g, const SNView& V, const PanelDesc& pd, int nrows, int ldx, double* xs) {
  for (int idx = g.rank(); idx < 3 * pd.w; idx += g.size())
    xs[idx * ldx + 3 * nrows] = sn_ld(V.z + 3 * static_cast<size_t>(pd.c0) + idx);
}

template <class G>
PGO_HD void sn_forward_fused(const G& g, const SNView& V, const PanelDesc& pd, int r0, int nrows,
                             int ldx, const double* xs, const double* Di, double* us, bool publish) {
  const int w = pd.w;
  for (int idx = g.rank(); idx < 3 * w; idx += g.size()) {
    const int t = idx / 3, k = idx % 3;
    const double* d = Di + 9 * t + 3 * k;
    const double v = d[0] * xs[(3 * t) * ldx + 3 * nrows] + d[1] * xs[(3 * t + 1) * ldx + 3 * nrows] +
                     d[2] * xs[(3 * t + 2) * ldx + 3 * nrows];
    us[idx] = v;
    if (publish) V.u[3 * static_cast<size_t>(pd.c0) + idx] = v;
  }
  g.sync();
  const int* rows = V.row_idx + pd.base + w + r0;
  for (int a = g.rank(); a < nrows; a += g.size()) {
    const int r = rows[a];
    double acc[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < 3 * w; ++j) {
      const double uj = us[j];
      acc[0] += xs[j * ldx + 3 * a] * uj;
      acc[1] += xs[j * ldx + 3 * a + 1] * uj;
      acc[2] += xs[j * ldx + 3 * a + 2] * uj;
    }
    for (int k = 0; k < 3; ++k) sn_add(V.z + 3 * static_cast<size_t>(r) + k, -acc[k]);
  }
}

// ---- factorisation tasks ---------------------------------------------------------------------------
// fa: factor the diagonal part (every row-chunk task of a panel repeats it: w <= 16, cheap) and
// finish the rows [r0, r1) below it. The r0 == 0 task publishes the diagonal part and Dinv.
template <class G>
PGO_HD void sn_task_factor(const G& g, const SNView& V, const Task& T, double* sm) {
  const PanelDesc pd = V.pn[T.id];
  const int w = pd.w, nrows = T.r1 - T.r0, ldx = 3 * nrows + 1;
  double* Dg = sm;
  double* Di = Dg + w * w * 9;
  int* pairs = reinterpret_cast<int*>(Di + w * 9);
  double* us = Di + w * 9 + kPairDoubles;
  double* xs = us + 3 * w;
  sn_load_rows_async(g, V, pd, T.r0, nrows, ldx, xs);  // in flight during the diagonal factorisation
  sn_load_diag(g, V, pd.base, w + pd.m, w, Dg);
  sn_load_rhs(g, V, pd, nrows, ldx, xs);
  g.sync();
  sn_factor_diag(g, V, w, Dg, Di, pairs);
  sn_copy_wait();
  g.sync();
  sn_solve_rows(g, w, Dg, xs, ldx, 3 * nrows + 1);
  g.sync();
  sn_store_rows(g, V, pd, T.r0, nrows, ldx, xs);
  sn_store_scaled_rows(g, V, pd, T.r0, nrows, ldx, xs, Di);
  if (T.r0 == 0) sn_store_diag(g, V, pd, Dg, Di, pd.scratch);
  sn_forward_fused(g, V, pd, T.r0, nrows, ldx, xs, Di, us, T.r0 == 0);
  g.sync();
}

PGO_HD int sn_target(const SNView& V, int meta, int a, int b) {
  return V.colbase[meta + b] + V.tbl[V.tbl_off[meta + b] + a];
}

// fb: one tile of an outer product
//   M(r_a, r_b) -= sum over the columns t of the panels [T.id - np + 1, T.id] of Y(a,t) M(b,t)^T,
//   Y(a,t) = M(a,t) Dinv_t,  a >= b rows of the LAST panel's below list.
// np = 1 and the "inside" flag: the update a panel owes the later columns of its own supernode
// (needed by the next panel); np = all panels of a supernode: the update the supernode owes its
// ancestors, applied once with the whole supernode as the inner dimension instead of once per
// panel (fewer, longer products and a fraction of the atomics).
// T.r0 = first row a, T.r1 = first column b, T.aux = ti | tj << 8 | np << 16 | inside << 28. The
// (0,0) tile also moves a scratch-published diagonal part into place (nothing reads it in this
// phase). This is the plain statement of the task, used by the host check; the device runs the same
// task as a tensor-core GEMM (sn_k_update in pgo_kernels.cu).
template <class G>
PGO_HD void sn_task_update(const G& g, const SNView& V, const Task& T, double* sm) {
  (void)sm;
  const int ti = T.aux & 0xFF, tj = (T.aux >> 8) & 0xFF, np = (T.aux >> 16) & 0x3F;
  const int q_last = T.id - ((T.aux >> 22) & 0x3F), q_first = q_last - np + 1;
  const bool inside = (T.aux >> 28) & 1;
  const PanelDesc last = V.pn[T.id];
  const int m = last.m, i0 = T.r0, j0 = T.r1;
  const int jlim = inside ? V.sn[last.sn].W - last.sn_off - last.w : m;
  const int ni = m - i0 < ti ? m - i0 : ti, nj = jlim - j0 < tj ? jlim - j0 : tj;
  if (i0 == 0 && j0 == 0 && q_last == T.id && last.scratch >= 0) {
    const int w = last.w, len = w + m;
    const double* src = V.scratch + 9 * static_cast<size_t>(last.scratch);
    for (int idx = g.rank(); idx < w * w * 9; idx += g.size()) {
      const int i = idx / (w * 9), t = (idx / 9) % w, k = idx % 9;
      if (i >= t) V.M[9 * static_cast<size_t>(sn_colpos(last.base, len, t) + (i - t)) + k] = sn_ld(src + idx);
    }
  }
  for (int item = g.rank(); item < ni * nj; item += g.size()) {
    const int a = i0 + item / nj, b = j0 + item % nj;
    if (a < b) continue;
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = q_first; q <= q_last; ++q) {
      const PanelDesc pd = V.pn[q];
      const int w = pd.w, len = w + pd.m, off = pd.m - m;
      for (int t = 0; t < w; ++t) {
        double mb[9], y[9];
        sn_ld9(V.Y + 9 * static_cast<size_t>(sn_colpos(pd.base, len, t) + (w - t) + off + a), y);
        sn_ld9(V.M + 9 * static_cast<size_t>(sn_colpos(pd.base, len, t) + (w - t) + off + b), mb);
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c)
            acc[3 * r + c] += y[3 * r] * mb[3 * c] + y[3 * r + 1] * mb[3 * c + 1] + y[3 * r + 2] * mb[3 * c + 2];
      }
    }
    double* dst = V.M + 9 * static_cast<size_t>(sn_target(V, last.meta, a, b));
    for (int k = 0; k < 9; ++k) sn_add(dst + k, -acc[k]);
  }
  g.sync();
}

// ff: a small panel (w <= kSmallWidth, at most kSmallRows rows below it) start to finish.
template <class G>
PGO_HD void sn_task_fused(const G& g, const SNView& V, const Task& T, double* sm) {
  const PanelDesc pd = V.pn[T.id];
  const int w = pd.w, m = pd.m, ldx = 3 * m + 1;
  double* Dg = sm;
  double* Di = Dg + w * w * 9;
  int* pairs = reinterpret_cast<int*>(Di + w * 9);
  double* us = Di + w * 9 + kSmallPairDoubles;
  double* xs = us + 3 * w;
  sn_load_rows_async(g, V, pd, 0, m, ldx, xs);  // in flight during the diagonal factorisation
  sn_load_diag(g, V, pd.base, w + m, w, Dg);
  sn_load_rhs(g, V, pd, m, ldx, xs);
  g.sync();
  sn_factor_diag(g, V, w, Dg, Di, pairs);
  sn_copy_wait();
  g.sync();
  sn_solve_rows(g, w, Dg, xs, ldx, 3 * m + 1);
  g.sync();
  sn_store_rows(g, V, pd, 0, m, ldx, xs);
  sn_store_diag(g, V, pd, Dg, Di, -1);
  sn_forward_fused(g, V, pd, 0, m, ldx, xs, Di, us, true);
  // outer product straight from shared memory: one (a, b) block pair, a >= b, per rank and turn
  // (pairs numbered column by column: column b holds the rows b .. m-1)
  for (int p = g.rank(); p < m * (m + 1) / 2; p += g.size()) {
    int b = 0, a = p;
    while (a >= m - b) {
      a -= m - b;
      ++b;
    }
    a += b;
    {
      const int tpos = V.colbase[pd.meta + b] + V.tbl[V.tbl_off[pd.meta + b] + a];
      double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int t = 0; t < w; ++t) {
        double av[9], y[9], bv[9];
        for (int k = 0; k < 9; ++k) {
          av[k] = xs[(3 * t + k % 3) * ldx + 3 * a + k / 3];
          bv[k] = xs[(3 * t + k % 3) * ldx + 3 * b + k / 3];
        }
        const double* d = Di + 9 * t;
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c)
            y[3 * r + c] = av[3 * r] * d[c] + av[3 * r + 1] * d[3 + c] + av[3 * r + 2] * d[6 + c];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c)
            acc[3 * r + c] += y[3 * r] * bv[3 * c] + y[3 * r + 1] * bv[3 * c + 1] + y[3 * r + 2] * bv[3 * c + 2];
      }
      double* dst = V.M + 9 * static_cast<size_t>(tpos);
      for (int k = 0; k < 9; ++k) sn_add(dst + k, -acc[k]);
    }
  }
  g.sync();
}

// ---- substitution tasks -----------------------------------------------------------------------------
PGO_HD void sn_mat_vec(const double* m, const double* v, double* out) {  // out = m v
  out[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  out[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  out[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}
PGO_HD void sn_mat_t_vec_add(const double* m, const double* v, double* acc) {  // acc += m^T v
  acc[0] += m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  acc[1] += m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  acc[2] += m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
}

template <class G>
PGO_HD void sn_load_dinv(const G& g, const SNView& V, int c0, int w, double* Di) {
  for (int idx = g.rank(); idx < 9 * w; idx += g.size())
    Di[idx] = sn_ld(V.Dinv + 9 * static_cast<size_t>(c0) + idx);
}

// ---- stand-alone forward substitution ---------------------------------------------------------------
// During an iteration the forward substitution rides along with the factorisation (sn_forward_fused).
// Further right-hand sides against the stored factor -- the unit block columns of the marginal
// covariances (SparseOptimizer::computeMarginals) -- go through these tasks, bottom-up
// over the supernode levels: sa / ss as in the backward direction, sf = rows below a wide
// supernode's panel (warp, chunks of 32 rows).
// Triangular part of one panel on staged vectors (all operands in shared memory).
template <class G>
PGO_HD void sn_forward_panel(const G& g, int w, const double* Dg, const double* Di, double* zs,
                             double* us) {
  for (int t = 0; t < w; ++t) {
    if (g.rank() == 0) sn_mat_vec(Di + 9 * t, zs + 3 * t, us + 3 * t);
    g.sync();
    for (int tp = t + 1 + g.rank(); tp < w; tp += g.size()) {
      double v[3];
      sn_mat_vec(Dg + (tp * w + t) * 9, us + 3 * t, v);
      zs[3 * tp] -= v[0];
      zs[3 * tp + 1] -= v[1];
      zs[3 * tp + 2] -= v[2];
    }
    g.sync();
  }
}

// rows [r0, r1) below the SUPERNODE of panel p:  z_{r_a} -= sum_t M(a,t) u_t  (atomic).
// us: the supernode's u in shared memory, or null to read it from global memory.
// Loads are issued four blocks at a time (memory-level parallelism, see sn_gather).
template <class G>
PGO_HD void sn_forward_rows(const G& g, const SNView& V, int p, int r0, int r1, const double* us) {
  const PanelDesc pd = V.pn[p];
  const SuperDesc sd = V.sn[pd.sn];
  const int w = pd.w, o = pd.sn_off, len = sd.W + sd.m;
  for (int a = r0 + g.rank(); a < r1; a += g.size()) {
    const int r = V.row_idx[sd.base + sd.W + a];
    double acc[3] = {0.0, 0.0, 0.0};
    for (int t0 = 0; t0 < w; t0 += 4) {
      double mb[4][9], uv[4][3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int q = 0; q < 4; ++q) {
        const int t = t0 + q;
        if (t < w) {
          sn_ld9(V.M + 9 * static_cast<size_t>(sn_colpos(sd.base, len, o + t) + (sd.W - o - t) + a), mb[q]);
          for (int k = 0; k < 3; ++k)
            uv[q][k] = us ? us[3 * (o + t) + k] : sn_ld(V.u + 3 * static_cast<size_t>(pd.c0 + t) + k);
        }
      }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int q = 0; q < 4; ++q)
        if (t0 + q < w) {
          double v[3];
          sn_mat_vec(mb[q], uv[q], v);
          acc[0] += v[0];
          acc[1] += v[1];
          acc[2] += v[2];
        }
    }
    for (int k = 0; k < 3; ++k) sn_add(V.z + 3 * static_cast<size_t>(r) + k, -acc[k]);
  }
}

// sa (forward): triangular part of a wide supernode, panel by panel, in one group.
// shared: zs[3W] us[3W] Dg[w*w*9] Di[w*9]
template <class G>
PGO_HD void sn_task_forward_tri(const G& g, const SNView& V, const Task& T, double* sm) {
  const SuperDesc sd = V.sn[T.id];
  const int W = sd.W, len = sd.W + sd.m;
  double* zs = sm;
  double* us = zs + 3 * W;
  double* Dg = us + 3 * W;
  double* Di = Dg + kPanelWidth * kPanelWidth * 9;
  sn_gather(g, 3 * W, [&](int idx, const double** sp, double** dp) {
    *sp = V.z + 3 * static_cast<size_t>(sd.c0) + idx;
    *dp = zs + idx;
  });
  for (int p = sd.pn_begin; p < sd.pn_end; ++p) {
    const PanelDesc pd = V.pn[p];
    const int w = pd.w, o = pd.sn_off;
    sn_load_diag(g, V, pd.base, w + pd.m, w, Dg);
    sn_load_dinv(g, V, pd.c0, w, Di);
    g.sync();
    sn_forward_panel(g, w, Dg, Di, zs + 3 * o, us + 3 * o);
    // the supernode's remaining columns are rows of this panel
    for (int tp = o + w + g.rank(); tp < W; tp += g.size()) {
      double acc[3] = {0.0, 0.0, 0.0};
      for (int t0 = 0; t0 < w; t0 += 4) {
        double mb[4][9];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 4; ++q)
          if (t0 + q < w)
            sn_ld9(V.M + 9 * static_cast<size_t>(sn_colpos(sd.base, len, o + t0 + q) + (tp - o - t0 - q)),
                   mb[q]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 4; ++q)
          if (t0 + q < w) {
            double v[3];
            sn_mat_vec(mb[q], us + 3 * (o + t0 + q), v);
            acc[0] += v[0];
            acc[1] += v[1];
            acc[2] += v[2];
          }
      }
      zs[3 * tp] -= acc[0];
      zs[3 * tp + 1] -= acc[1];
      zs[3 * tp + 2] -= acc[2];
    }
    g.sync();
  }
  for (int idx = g.rank(); idx < 3 * W; idx += g.size()) V.u[3 * static_cast<size_t>(sd.c0) + idx] = us[idx];
  g.sync();
}

// ss (forward): a narrow supernode (one panel) including the rows below it.
// shared: zs[3W] us[3W] Dg[W*W*9] Di[W*9]
template <class G>
PGO_HD void sn_task_forward_small(const G& g, const SNView& V, const Task& T, double* sm) {
  const SuperDesc sd = V.sn[T.id];
  const int W = sd.W;
  double* zs = sm;
  double* us = zs + 3 * W;
  double* Dg = us + 3 * W;
  double* Di = Dg + W * W * 9;
  for (int idx = g.rank(); idx < 3 * W; idx += g.size())
    zs[idx] = sn_ld(V.z + 3 * static_cast<size_t>(sd.c0) + idx);
  sn_load_diag(g, V, sd.base, W + sd.m, W, Dg);
  sn_load_dinv(g, V, sd.c0, W, Di);
  g.sync();
  sn_forward_panel(g, W, Dg, Di, zs, us);
  for (int idx = g.rank(); idx < 3 * W; idx += g.size()) V.u[3 * static_cast<size_t>(sd.c0) + idx] = us[idx];
  sn_forward_rows(g, V, sd.pn_begin, 0, sd.m, us);
  g.sync();
}

// acc += sum over the rows a = first, first + step, ... < end of  M(a, col)^T x_{row(a)}, where the
// blocks of the column are consecutive from `col` and xv(a, out) fetches x of row a. Four rows in
// flight at a time.
template <class XF>
PGO_HD void sn_col_dot(const SNView& V, size_t col, int first, int step, int end, XF xv, double* acc) {
  for (int a0 = first; a0 < end; a0 += 4 * step) {
    double mb[4][9], x[4][3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 4; ++q) {
      const int a = a0 + q * step;
      if (a < end) {
        xv(a, x[q]);
        sn_ld9(V.M + 9 * (col + a), mb[q]);
      }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 4; ++q)
      if (a0 + q * step < end) sn_mat_t_vec_add(mb[q], x[q], acc);
  }
}

// Backward, rows [r0, r1) below the supernode of panel p:  x_{c0+t} += sum_a M(a,t)^T x_{r_a}
// (x doubles as the accumulator of a wide supernode's columns; atomic).
template <class G>
PGO_HD void sn_backward_rows(const G& g, const SNView& V, int p, int r0, int r1) {
  const PanelDesc pd = V.pn[p];
  const SuperDesc sd = V.sn[pd.sn];
  const int w = pd.w, o = pd.sn_off, len = sd.W + sd.m;
  const int split = g.size() / w > 0 ? g.size() / w : 1;
  const int* rows = V.row_idx + sd.base + sd.W;
  for (int item = g.rank(); item < w * split; item += g.size()) {
    const int t = item % w, h = item / w;
    const size_t col = static_cast<size_t>(sn_colpos(sd.base, len, o + t) + (sd.W - o - t));
    double acc[3] = {0.0, 0.0, 0.0};
    sn_col_dot(V, col, r0 + h, split, r1, [&](int a, double* out) {
      const int r = rows[a];
      for (int k = 0; k < 3; ++k) out[k] = sn_ld(V.x + 3 * static_cast<size_t>(r) + k);
    }, acc);
    for (int k = 0; k < 3; ++k) sn_add(V.x + 3 * static_cast<size_t>(pd.c0 + t) + k, acc[k]);
  }
}

// Triangular part of one panel, backward: xs holds the accumulators on entry, x on return.
template <class G>
PGO_HD void sn_backward_panel(const G& g, int w, const double* Dg, const double* Di, const double* us,
                              double* xs) {
  for (int t = w - 1; t >= 0; --t) {
    if (g.rank() == 0) {
      double v[3];
      sn_mat_vec(Di + 9 * t, xs + 3 * t, v);
      for (int k = 0; k < 3; ++k) xs[3 * t + k] = us[3 * t + k] - v[k];
    }
    g.sync();
    for (int sidx = g.rank(); sidx < t; sidx += g.size())
      sn_mat_t_vec_add(Dg + (t * w + sidx) * 9, xs + 3 * t, xs + 3 * sidx);  // M(t,s)^T x_t
    g.sync();
  }
}

// sa (backward): triangular part of a wide supernode, last panel first.
// shared: xs[3W] us[3W] Dg[w*w*9] Di[w*9] red[3 * max(size, w)]
template <class G>
PGO_HD void sn_task_backward_tri(const G& g, const SNView& V, const Task& T, double* sm) {
  const SuperDesc sd = V.sn[T.id];
  const int W = sd.W, len = sd.W + sd.m;
  double* xs = sm;
  double* us = xs + 3 * W;
  double* Dg = us + 3 * W;
  double* Di = Dg + kPanelWidth * kPanelWidth * 9;
  double* red = Di + kPanelWidth * 9;
  sn_gather(g, 6 * W, [&](int idx, const double** sp, double** dp) {
    const int j = idx < 3 * W ? idx : idx - 3 * W;
    *sp = (idx < 3 * W ? V.x : V.u) + 3 * static_cast<size_t>(sd.c0) + j;
    *dp = (idx < 3 * W ? xs : us) + j;
  });
  for (int p = sd.pn_end - 1; p >= sd.pn_begin; --p) {
    const PanelDesc pd = V.pn[p];
    const int w = pd.w, o = pd.sn_off;
    sn_load_diag(g, V, pd.base, w + pd.m, w, Dg);
    sn_load_dinv(g, V, pd.c0, w, Di);
    g.sync();  // xs of the later panels is final
    // contributions of the supernode's later columns to this panel's columns
    const int split = g.size() / w > 0 ? g.size() / w : 1;
    for (int item = g.rank(); item < w * split; item += g.size()) {
      const int t = item % w, h = item / w;
      // block (tp, o + t) sits at colpos(o + t) + (tp - o - t): index the column by tp directly
      const size_t col = static_cast<size_t>(sn_colpos(sd.base, len, o + t)) - (o + t);
      double acc[3] = {0.0, 0.0, 0.0};
      sn_col_dot(V, col, o + w + h, split, W, [&](int tp, double* out) {
        for (int k = 0; k < 3; ++k) out[k] = xs[3 * tp + k];
      }, acc);
      for (int k = 0; k < 3; ++k) red[3 * item + k] = acc[k];
    }
    g.sync();
    for (int idx = g.rank(); idx < 3 * w; idx += g.size()) {
      const int t = idx / 3, k = idx % 3;
      double sum = 0.0;
      for (int h = 0; h < split; ++h) sum += red[3 * (h * w + t) + k];
      xs[3 * (o + t) + k] += sum;
    }
    g.sync();
    sn_backward_panel(g, w, Dg, Di, us + 3 * o, xs + 3 * o);
  }
  for (int idx = g.rank(); idx < 3 * W; idx += g.size()) V.x[3 * static_cast<size_t>(sd.c0) + idx] = xs[idx];
  g.sync();
}

// ss (backward): a narrow supernode including the gather over the rows below it.
// shared: xs[3W] us[3W] Dg[W*W*9] Di[W*9] red[3 * max(size, W)]
template <class G>
PGO_HD void sn_task_backward_small(const G& g, const SNView& V, const Task& T, double* sm) {
  const SuperDesc sd = V.sn[T.id];
  const int W = sd.W, m = sd.m, len = W + m;
  double* xs = sm;
  double* us = xs + 3 * W;
  double* Dg = us + 3 * W;
  double* Di = Dg + W * W * 9;
  double* red = Di + W * 9;
  sn_load_diag(g, V, sd.base, len, W, Dg);
  sn_load_dinv(g, V, sd.c0, W, Di);
  for (int idx = g.rank(); idx < 3 * W; idx += g.size())
    us[idx] = sn_ld(V.u + 3 * static_cast<size_t>(sd.c0) + idx);
  const int split = g.size() / W > 0 ? g.size() / W : 1;
  const int* rows = V.row_idx + sd.base + W;
  for (int item = g.rank(); item < W * split; item += g.size()) {
    const int t = item % W, h = item / W;
    const size_t col = static_cast<size_t>(sn_colpos(sd.base, len, t) + (W - t));
    double acc[3] = {0.0, 0.0, 0.0};
    sn_col_dot(V, col, h, split, m, [&](int a, double* out) {
      const int r = rows[a];
      for (int k = 0; k < 3; ++k) out[k] = sn_ld(V.x + 3 * static_cast<size_t>(r) + k);
    }, acc);
    for (int k = 0; k < 3; ++k) red[3 * item + k] = acc[k];
  }
  g.sync();
  for (int idx = g.rank(); idx < 3 * W; idx += g.size()) {
    const int t = idx / 3, k = idx % 3;
    double sum = 0.0;
    for (int h = 0; h < split; ++h) sum += red[3 * (h * W + t) + k];
    xs[idx] = sum;
  }
  g.sync();
  sn_backward_panel(g, W, Dg, Di, us, xs);
  for (int idx = g.rank(); idx < 3 * W; idx += g.size()) V.x[3 * static_cast<size_t>(sd.c0) + idx] = xs[idx];
  g.sync();
}

}  // namespace pgo
#endif
