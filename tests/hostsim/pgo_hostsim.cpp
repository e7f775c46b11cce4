// TEST INFRASTRUCTURE ONLY -- never linked into libcgmrslam_b200.so, never loaded by the package.
//
// Runs the supernodal factorisation / substitution TASK FUNCTIONS of the product
// (cg_mrslam_b200/csrc/pgo_supernodal.h -- the same source the CUDA kernels instantiate with CTA
// and warp groups) with a one-thread group on the host, driven by the product's own structure
// analysis (pgo_symbolic.cpp). This checks the supernode / panel tables, the scatter tables, the
// task lists and the task arithmetic on a GPU-less box (`pytest -m "not gpu"`) against scipy.
// It is not a fallback: the package cannot reach it.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "pgo_supernodal.h"

using namespace pgo;

// Solve H x = b for a block matrix given as lower-triangular block triplets (row >= col in free-
// vertex numbering, 3x3 row-major, duplicates add). Returns 0, 1 (analysis failed), 2 (not SPD).
extern "C" int pgo_hostsim_solve(int n, int n_blocks, const int32_t* brow, const int32_t* bcol,
                                 const double* bval, const double* rhs, double* x_out,
                                 int32_t* stats /*[8]*/) {
  std::vector<std::pair<int, int> > edges;
  for (int k = 0; k < n_blocks; ++k)
    if (brow[k] != bcol[k]) edges.push_back(std::make_pair(bcol[k], brow[k]));
  Symbolic S;
  std::string err;
  if (!analyse(n, edges, 0, 1, &S, &err)) return 1;
  const Supernodal& N = S.sn;
  std::vector<double> M(9 * static_cast<size_t>(S.nnzb), 0.0), Dinv(9 * static_cast<size_t>(n)),
      z(3 * static_cast<size_t>(n)), u(3 * static_cast<size_t>(n)), x(3 * static_cast<size_t>(n), 0.0),
      scratch(9 * static_cast<size_t>(N.scratch_blocks) + 9);
  for (int k = 0; k < n_blocks; ++k) {
    int pr = S.iperm[brow[k]], pc = S.iperm[bcol[k]];
    double v[9];
    std::memcpy(v, bval + 9 * static_cast<size_t>(k), sizeof v);
    if (pr < pc) {  // transpose into the lower triangle of the permuted matrix
      std::swap(pr, pc);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) v[3 * r + c] = bval[9 * static_cast<size_t>(k) + 3 * c + r];
    }
    int pos = -1;
    for (int w = S.col_ptr[pc]; w < S.col_ptr[pc + 1]; ++w)
      if (S.row_idx[w] == pr) pos = w;
    if (pos < 0) return 1;
    for (int i = 0; i < 9; ++i) M[9 * static_cast<size_t>(pos) + i] += v[i];
  }
  for (int p = 0; p < n; ++p)
    for (int k = 0; k < 3; ++k) z[3 * static_cast<size_t>(p) + k] = rhs[3 * static_cast<size_t>(S.perm[p]) + k];
  int status = 0;
  SNView V;
  V.row_idx = S.row_idx.data();
  V.pn = N.pn.data();
  V.sn = N.sn.data();
  V.colbase = N.colbase.data();
  V.tbl_off = N.tbl_off.data();
  V.tbl = N.tbl.data();
  std::vector<double> Y(M.size(), 0.0);
  V.M = M.data();
  V.Y = Y.data();
  V.Dinv = Dinv.data();
  V.z = z.data();
  V.u = u.data();
  V.x = x.data();
  V.scratch = scratch.data();
  V.status = &status;
  V.s_M = V.s_Dinv = V.s_vec = V.s_scratch = 0;
  V.s_status = 0;
  std::vector<double> sm(kCtaSmemDoubles);
  const SeqGroup g;
  for (int l = 0; l < N.n_plevels; ++l) {
    for (int i = N.fa_ptr[l]; i < N.fa_ptr[l + 1]; ++i) sn_task_factor(g, V, N.fa[i], sm.data());
    for (int i = N.ff_ptr[l]; i < N.ff_ptr[l + 1]; ++i) sn_task_fused(g, V, N.ff[i], sm.data());
    for (int i = N.fb_ptr[l]; i < N.fb_ptr[l + 1]; ++i) sn_task_update(g, V, N.fb[i], sm.data());
  }
  if (status) return 2;
  for (int l = N.n_slevels - 1; l >= 0; --l) {
    for (int i = N.sb_ptr[l]; i < N.sb_ptr[l + 1]; ++i)
      sn_backward_rows(g, V, N.sb[i].id, N.sb[i].r0, N.sb[i].r1);
    for (int i = N.ss_ptr[l]; i < N.ss_ptr[l + 1]; ++i) sn_task_backward_small(g, V, N.ss[i], sm.data());
    for (int i = N.sa_ptr[l]; i < N.sa_ptr[l + 1]; ++i) sn_task_backward_tri(g, V, N.sa[i], sm.data());
  }
  for (int p = 0; p < n; ++p)
    for (int k = 0; k < 3; ++k) x_out[3 * static_cast<size_t>(S.perm[p]) + k] = x[3 * static_cast<size_t>(p) + k];
  // the stand-alone forward substitution (marginals path) on the same right-hand side must give the
  // same solution: x2 is returned behind x
  {
    std::vector<double> z2(3 * static_cast<size_t>(n)), x2(3 * static_cast<size_t>(n), 0.0);
    for (int p = 0; p < n; ++p)
      for (int k = 0; k < 3; ++k) z2[3 * static_cast<size_t>(p) + k] = rhs[3 * static_cast<size_t>(S.perm[p]) + k];
    V.z = z2.data();
    V.x = x2.data();
    for (int l = 0; l < N.n_slevels; ++l) {
      for (int i = N.sa_ptr[l]; i < N.sa_ptr[l + 1]; ++i) sn_task_forward_tri(g, V, N.sa[i], sm.data());
      for (int i = N.ss_ptr[l]; i < N.ss_ptr[l + 1]; ++i) sn_task_forward_small(g, V, N.ss[i], sm.data());
      for (int i = N.sf_ptr[l]; i < N.sf_ptr[l + 1]; ++i)
        sn_forward_rows(g, V, N.sf[i].id, N.sf[i].r0, N.sf[i].r1, nullptr);
    }
    for (int l = N.n_slevels - 1; l >= 0; --l) {
      for (int i = N.sb_ptr[l]; i < N.sb_ptr[l + 1]; ++i) sn_backward_rows(g, V, N.sb[i].id, N.sb[i].r0, N.sb[i].r1);
      for (int i = N.ss_ptr[l]; i < N.ss_ptr[l + 1]; ++i) sn_task_backward_small(g, V, N.ss[i], sm.data());
      for (int i = N.sa_ptr[l]; i < N.sa_ptr[l + 1]; ++i) sn_task_backward_tri(g, V, N.sa[i], sm.data());
    }
    for (int p = 0; p < n; ++p)
      for (int k = 0; k < 3; ++k)
        x_out[3 * static_cast<size_t>(n) + 3 * static_cast<size_t>(S.perm[p]) + k] = x2[3 * static_cast<size_t>(p) + k];
  }
  if (stats) {
    stats[0] = N.n_super;
    stats[1] = N.n_panels;
    stats[2] = N.n_plevels;
    stats[3] = N.n_slevels;
    stats[4] = static_cast<int32_t>(N.fa.size());
    stats[5] = static_cast<int32_t>(N.fb.size());
    stats[6] = static_cast<int32_t>(N.ff.size());
    stats[7] = static_cast<int32_t>(N.scratch_blocks);
  }
  return 0;
}

// Structure analysis alone: out = {factor blocks, MFLOP of one factorisation, panel levels,
// supernode levels, supernodes, shared (top separator) columns}. Returns 0, or 1 with the analysis
// refusing the graph (its own checks: every edge inside one owner or touching a shared separator,
// descendants of a column owned like the column).
extern "C" int pgo_hostsim_analyse(int n, int n_edges, const int32_t* ei, const int32_t* ej, int world,
                                   int64_t* out /*[6]*/) {
  std::vector<std::pair<int, int> > edges(n_edges);
  for (int k = 0; k < n_edges; ++k) edges[k] = std::make_pair(ei[k], ej[k]);
  Symbolic S;
  std::string err;
  if (!analyse(n, edges, 0, world, &S, &err)) return 1;
  out[0] = S.nnzb;
  out[1] = static_cast<int64_t>(S.sn.flops / 1e6);
  out[2] = S.sn.n_plevels;
  out[3] = S.sn.n_slevels;
  out[4] = S.sn.n_super;
  out[5] = S.n - S.first_shared;
  return 0;
}

