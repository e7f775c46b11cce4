// TEST INFRASTRUCTURE ONLY -- compiled, single-thread CPU restatement of the pose-graph solve the
// reference delegates to g2o, organised the way g2o's own stack runs it:
//   EdgeSE2::computeError / linearizeOplus / constructQuadraticForm per edge  
//   BlockSolver::buildSystem into an upper block-triangular H                 (C5, C6)
//   LinearSolverCSparse::solve with setBlockOrdering(false): scalar CCS of the upper triangle,
//   a fill-reducing ordering, sparse up-looking Cholesky, two triangular solves (C8; configured by
//   the reference at src/slam/graph_slam.cpp:44-56)
//   OptimizationAlgorithmGaussNewton: exactly n iterations, additive oplus     (C3, C7)
//
// PARITY UNPINNED, like oracle/pgo_oracle.py: g2o and CSparse are third-party, absent from
// /root/reference and from this container. This file restates the published algorithms
// (up-looking sparse Cholesky with elimination-tree reach, as in T. Davis, "Direct Methods for
// Sparse Linear Systems", ch. 4; a quotient-graph minimum-degree ordering with exact external
// degrees standing in for CSparse's AMD -- the ordering changes rounding and fill, not the
// solution). Two uses, both outside the product:
//   * tests: a third, independent solve (up-looking scalar Cholesky) next to scipy's SuperLU
//     (pgo_oracle.py) and the product's supernodal block LDL^T;
//   * bench.py cpu_baseline / --impl reference: the fair single-thread CPU arm for the GN metric
//     (the Python oracle spends most of its time in numpy glue).
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <set>
#include <string>
#include <utility>
#include <vector>

namespace {

const double kPi = 3.14159265358979323846;

double normalize_theta(double t) {  // C2 (older g2o form)
  if (t >= -kPi && t < kPi) return t;
  const double two_pi = 2.0 * kPi;
  double w = t - two_pi * std::floor(t / two_pi);
  if (w >= kPi) w -= two_pi;
  return w;
}

struct SE2 {
  double x, y, th;
};
SE2 inv(const SE2& a) {  // C1
  SE2 r;
  r.th = normalize_theta(-a.th);
  const double c = std::cos(r.th), s = std::sin(r.th);
  r.x = c * (-a.x) - s * (-a.y);
  r.y = s * (-a.x) + c * (-a.y);
  return r;
}
SE2 mul(const SE2& a, const SE2& b) {  // C1
  const double c = std::cos(a.th), s = std::sin(a.th);
  SE2 r;
  r.x = a.x + (c * b.x - s * b.y);
  r.y = a.y + (s * b.x + c * b.y);
  r.th = normalize_theta(a.th + b.th);
  return r;
}

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- minimum-degree ordering of a symmetric block pattern (quotient graph) ------------------------
// adj: neighbours of every node (no self loops, symmetric). Returns the elimination order.
std::vector<int> minimum_degree(int n, const std::vector<std::vector<int> >& adj) {
  std::vector<std::vector<int> > A(adj);         // variable neighbours still uneliminated
  std::vector<std::vector<int> > E(n);           // elements adjacent to a variable
  std::vector<std::vector<int> > L(n);           // variables of an element
  std::vector<char> dead(n, 0), absorbed(n, 0);  // eliminated variable / absorbed element
  std::vector<int> mark(n, -1), deg(n);
  std::set<std::pair<int, int> > heap;
  for (int i = 0; i < n; ++i) {
    deg[i] = static_cast<int>(A[i].size());
    heap.insert(std::make_pair(deg[i], i));
  }
  std::vector<int> order;
  order.reserve(n);
  int stamp = 0;
  std::vector<int> Lp;
  for (int step = 0; step < n; ++step) {
    const int p = heap.begin()->second;
    heap.erase(heap.begin());
    order.push_back(p);
    dead[p] = 1;
    // Lp = A_p  U  union of L_e over e in E_p, minus p
    ++stamp;
    mark[p] = stamp;
    Lp.clear();
    for (size_t k = 0; k < A[p].size(); ++k) {
      const int v = A[p][k];
      if (!dead[v] && mark[v] != stamp) {
        mark[v] = stamp;
        Lp.push_back(v);
      }
    }
    for (size_t k = 0; k < E[p].size(); ++k) {
      const int e = E[p][k];
      if (absorbed[e]) continue;
      for (size_t q = 0; q < L[e].size(); ++q) {
        const int v = L[e][q];
        if (!dead[v] && mark[v] != stamp) {
          mark[v] = stamp;
          Lp.push_back(v);
        }
      }
      absorbed[e] = 1;  // e is absorbed into the new element p
      std::vector<int>().swap(L[e]);
    }
    const int lp_stamp = stamp;
    L[p] = Lp;
    std::vector<int>().swap(A[p]);
    std::vector<int>().swap(E[p]);
    for (size_t t = 0; t < Lp.size(); ++t) {
      const int i = Lp[t];
      heap.erase(std::make_pair(deg[i], i));
      // prune: variables now covered by element p, dead ones; absorbed elements
      size_t w = 0;
      for (size_t k = 0; k < A[i].size(); ++k) {
        const int v = A[i][k];
        if (!dead[v] && mark[v] != lp_stamp) A[i][w++] = v;
      }
      A[i].resize(w);
      w = 0;
      for (size_t k = 0; k < E[i].size(); ++k)
        if (!absorbed[E[i][k]]) E[i][w++] = E[i][k];
      E[i].resize(w);
      E[i].push_back(p);
    }
    // exact external degrees of the touched variables
    for (size_t t = 0; t < Lp.size(); ++t) {
      const int i = Lp[t];
      ++stamp;
      mark[i] = stamp;
      int d = 0;
      for (size_t k = 0; k < A[i].size(); ++k) {
        const int v = A[i][k];
        if (mark[v] != stamp) {
          mark[v] = stamp;
          ++d;
        }
      }
      for (size_t k = 0; k < E[i].size(); ++k) {
        const std::vector<int>& le = L[E[i][k]];
        for (size_t q = 0; q < le.size(); ++q) {
          const int v = le[q];
          if (!dead[v] && mark[v] != stamp) {
            mark[v] = stamp;
            ++d;
          }
        }
      }
      deg[i] = d;
      heap.insert(std::make_pair(d, i));
    }
    // restore the Lp marks for nothing: stamps only grow, so no reset is needed
  }
  return order;
}

// ---- sparse symmetric matrix, upper triangle in CCS, and its up-looking Cholesky --------------------
struct Chol {
  int n = 0;
  std::vector<int> perm, iperm;    // perm[k] = original scalar index eliminated k-th
  std::vector<int> parent;         // elimination tree of the permuted matrix
  std::vector<int> Cp, Ci;         // permuted upper triangle C = P A P^T, CCS, pattern only here
  std::vector<int> amap;           // for every entry of C: index into the values of A's upper CCS
  std::vector<int> Lp, Li;         // factor, CCS, diagonal first in every column
  std::vector<double> Lx;
  long long flops = 0;
};

// nonzero pattern of row k of L: the nodes reached from the entries of column k of C (rows < k) up
// the elimination tree; returned in topological order in s[top .. n)
int ereach(const Chol& F, int k, std::vector<int>& s, std::vector<int>& w) {
  int top = F.n;
  w[k] = k;
  for (int p = F.Cp[k]; p < F.Cp[k + 1]; ++p) {
    int i = F.Ci[p];
    if (i > k) continue;
    int len = 0;
    for (; w[i] != k; i = F.parent[i]) {
      s[len++] = i;
      w[i] = k;
    }
    while (len > 0) s[--top] = s[--len];
  }
  return top;
}

// Ap/Ai: upper triangle (row <= column) of the scalar matrix, CCS; block_order: elimination order of
// the 3x3 block variables.
void symbolic(int n, const std::vector<int>& Ap, const std::vector<int>& Ai,
              const std::vector<int>& block_order, Chol* F) {
  F->n = n;
  F->perm.resize(n);
  F->iperm.resize(n);
  for (size_t k = 0; k < block_order.size(); ++k)
    for (int c = 0; c < 3; ++c) {
      F->perm[3 * k + c] = 3 * block_order[k] + c;
      F->iperm[3 * block_order[k] + c] = static_cast<int>(3 * k + c);
    }
  // C = P A P^T, upper triangle
  std::vector<int> cnt(n + 1, 0);
  for (int j = 0; j < n; ++j)
    for (int p = Ap[j]; p < Ap[j + 1]; ++p) {
      const int i2 = F->iperm[Ai[p]], j2 = F->iperm[j];
      ++cnt[std::max(i2, j2) + 1];
    }
  F->Cp.assign(n + 1, 0);
  for (int j = 0; j < n; ++j) F->Cp[j + 1] = F->Cp[j] + cnt[j + 1];
  F->Ci.resize(F->Cp[n]);
  F->amap.resize(F->Cp[n]);
  std::vector<int> cur(F->Cp.begin(), F->Cp.end() - 1);
  for (int j = 0; j < n; ++j)
    for (int p = Ap[j]; p < Ap[j + 1]; ++p) {
      const int i2 = F->iperm[Ai[p]], j2 = F->iperm[j];
      const int q = cur[std::max(i2, j2)]++;
      F->Ci[q] = std::min(i2, j2);
      F->amap[q] = p;
    }
  // elimination tree (path compression with ancestors)
  F->parent.assign(n, -1);
  std::vector<int> anc(n, -1);
  for (int k = 0; k < n; ++k)
    for (int p = F->Cp[k]; p < F->Cp[k + 1]; ++p) {
      int i = F->Ci[p];
      while (i != -1 && i < k) {
        const int next = anc[i];
        anc[i] = k;
        if (next == -1) F->parent[i] = k;
        i = next;
      }
    }
  // column counts by row reach (O(nnz(L)))
  std::vector<int> s(n), w(n, -1), colcount(n, 1);
  for (int k = 0; k < n; ++k) {
    const int top = ereach(*F, k, s, w);
    for (int t = top; t < n; ++t) ++colcount[s[t]];
  }
  F->Lp.assign(n + 1, 0);
  long long fl = 0;
  for (int j = 0; j < n; ++j) {
    F->Lp[j + 1] = F->Lp[j] + colcount[j];
    fl += static_cast<long long>(colcount[j]) * colcount[j];
  }
  F->flops = fl;
  F->Li.resize(F->Lp[n]);
  F->Lx.resize(F->Lp[n]);
}

// numeric up-looking Cholesky; Ax = values of A's upper CCS. false if not positive definite.
bool numeric(const std::vector<double>& Ax, Chol* F) {
  const int n = F->n;
  std::vector<int> c(F->Lp.begin(), F->Lp.end() - 1), s(n), w(n, -1);
  std::vector<double> x(n, 0.0);
  for (int k = 0; k < n; ++k) {
    const int top = ereach(*F, k, s, w);
    x[k] = 0.0;
    for (int p = F->Cp[k]; p < F->Cp[k + 1]; ++p)
      if (F->Ci[p] <= k) x[F->Ci[p]] = Ax[F->amap[p]];
    double d = x[k];
    x[k] = 0.0;
    for (int t = top; t < n; ++t) {
      const int i = s[t];
      const double lki = x[i] / F->Lx[F->Lp[i]];
      x[i] = 0.0;
      for (int p = F->Lp[i] + 1; p < c[i]; ++p) x[F->Li[p]] -= F->Lx[p] * lki;
      d -= lki * lki;
      const int p = c[i]++;
      F->Li[p] = k;
      F->Lx[p] = lki;
    }
    if (!(d > 0.0)) return false;
    const int p = c[k]++;
    F->Li[p] = k;
    F->Lx[p] = std::sqrt(d);
  }
  return true;
}

void solve(const Chol& F, const std::vector<double>& b, std::vector<double>* out) {
  const int n = F.n;
  std::vector<double> y(n);
  for (int k = 0; k < n; ++k) y[k] = b[F.perm[k]];
  for (int j = 0; j < n; ++j) {  // L y = Pb
    y[j] /= F.Lx[F.Lp[j]];
    for (int p = F.Lp[j] + 1; p < F.Lp[j + 1]; ++p) y[F.Li[p]] -= F.Lx[p] * y[j];
  }
  for (int j = n - 1; j >= 0; --j) {  // L^T z = y
    for (int p = F.Lp[j] + 1; p < F.Lp[j + 1]; ++p) y[j] -= F.Lx[p] * y[F.Li[p]];
    y[j] /= F.Lx[F.Lp[j]];
  }
  out->resize(n);
  for (int k = 0; k < n; ++k) (*out)[F.perm[k]] = y[k];
}

// ---- one pose graph: structure, numeric state, the Gauss-Newton iteration ---------------------------
struct Graph {
  int nv = 0, ne = 0, nb = 0, n = 0;
  std::vector<int> ei, ej, hidx;
  std::vector<uint8_t> fixed;
  std::vector<std::vector<int> > up;  // block rows of block column j (ascending, diagonal last)
  std::vector<int> Ap, Ai;
  std::vector<double> Ax, b;
  Chol F;
  bool have_factor = false;
  double t_analyse = 0.0, t_lin = 0.0, t_factor = 0.0, t_solve = 0.0;

  double* entry(int i, int j, int r, int c) {  // block (i, j), i <= j, entry (r, c)
    const size_t k = std::lower_bound(up[j].begin(), up[j].end(), i) - up[j].begin();
    return &Ax[Ap[3 * j + c] + 3 * static_cast<int>(k) + r];
  }

  void set_structure(int nv_, int ne_, const int32_t* ei_, const int32_t* ej_, const uint8_t* fixed_) {
    const double t0 = now();
    nv = nv_;
    ne = ne_;
    ei.assign(ei_, ei_ + ne);
    ej.assign(ej_, ej_ + ne);
    fixed.assign(fixed_, fixed_ + nv);
    hidx.assign(nv, -1);  // non-fixed vertices in vertex order (C6)
    nb = 0;
    for (int v = 0; v < nv; ++v)
      if (!fixed[v]) hidx[v] = nb++;
    n = 3 * nb;
    std::vector<std::vector<int> > adj(nb);
    for (int e = 0; e < ne; ++e) {
      const int a = hidx[ei[e]], c = hidx[ej[e]];
      if (a >= 0 && c >= 0 && a != c) {
        adj[a].push_back(c);
        adj[c].push_back(a);
      }
    }
    for (int v = 0; v < nb; ++v) {
      std::sort(adj[v].begin(), adj[v].end());
      adj[v].erase(std::unique(adj[v].begin(), adj[v].end()), adj[v].end());
    }
    // scalar CCS of the upper triangle: block column j holds the blocks (i, j), i <= j
    up.assign(nb, std::vector<int>());
    for (int j = 0; j < nb; ++j) {
      for (size_t k = 0; k < adj[j].size(); ++k)
        if (adj[j][k] < j) up[j].push_back(adj[j][k]);
      up[j].push_back(j);
    }
    Ap.assign(n + 1, 0);
    Ai.clear();
    for (int j = 0; j < nb; ++j)
      for (int c = 0; c < 3; ++c) {
        for (size_t k = 0; k < up[j].size(); ++k) {
          const int i = up[j][k];
          for (int r = 0; r < 3; ++r)
            if (i < j || r <= c) Ai.push_back(3 * i + r);
        }
        Ap[3 * j + c + 1] = static_cast<int>(Ai.size());
      }
    Ax.assign(Ai.size(), 0.0);
    b.assign(n, 0.0);
    F = Chol();
    symbolic(n, Ap, Ai, minimum_degree(nb, adj), &F);
    have_factor = false;
    t_analyse = now() - t0;
  }

  // C4 + C5 at `poses`; returns chi2
  double linearise(const double* poses, const double* meas, const double* info6) {
    std::fill(Ax.begin(), Ax.end(), 0.0);
    std::fill(b.begin(), b.end(), 0.0);
    double chi2 = 0.0;
    for (int e = 0; e < ne; ++e) {
      const int vi = ei[e], vj = ej[e];
      const SE2 xi = {poses[3 * vi], poses[3 * vi + 1], poses[3 * vi + 2]};
      const SE2 xj = {poses[3 * vj], poses[3 * vj + 1], poses[3 * vj + 2]};
      const SE2 z = {meas[3 * e], meas[3 * e + 1], meas[3 * e + 2]};
      const SE2 zi = inv(z);
      const SE2 er = mul(zi, mul(inv(xi), xj));  // C4
      const double err[3] = {er.x, er.y, er.th};
      const double ci = std::cos(xi.th), si = std::sin(xi.th);
      const double ddx = xj.x - xi.x, ddy = xj.y - xi.y;
      const double a[9] = {-ci, -si, -si * ddx + ci * ddy, si, -ci, -ci * ddx - si * ddy, 0, 0, -1};
      const double bj[9] = {ci, si, 0, -si, ci, 0, 0, 0, 1};
      const double cz = std::cos(zi.th), sz = std::sin(zi.th);
      double Ji[9], Jj[9];
      for (int c = 0; c < 3; ++c) {
        Ji[c] = cz * a[c] - sz * a[3 + c];
        Ji[3 + c] = sz * a[c] + cz * a[3 + c];
        Ji[6 + c] = a[6 + c];
        Jj[c] = cz * bj[c] - sz * bj[3 + c];
        Jj[3 + c] = sz * bj[c] + cz * bj[3 + c];
        Jj[6 + c] = bj[6 + c];
      }
      const double* w = info6 + 6 * static_cast<size_t>(e);
      const double om[9] = {w[0], w[1], w[2], w[1], w[3], w[4], w[2], w[4], w[5]};
      double oe[3];
      for (int r = 0; r < 3; ++r) oe[r] = om[3 * r] * err[0] + om[3 * r + 1] * err[1] + om[3 * r + 2] * err[2];
      chi2 += err[0] * oe[0] + err[1] * oe[1] + err[2] * oe[2];
      const int hi = hidx[vi], hj = hidx[vj];
      double AiT[9], AjT[9];  // J^T Omega
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          AiT[3 * r + c] = Ji[r] * om[c] + Ji[3 + r] * om[3 + c] + Ji[6 + r] * om[6 + c];
          AjT[3 * r + c] = Jj[r] * om[c] + Jj[3 + r] * om[3 + c] + Jj[6 + r] * om[6 + c];
        }
      if (hi >= 0)
        for (int r = 0; r < 3; ++r) {
          b[3 * hi + r] -= AiT[3 * r] * err[0] + AiT[3 * r + 1] * err[1] + AiT[3 * r + 2] * err[2];
          for (int c = r; c < 3; ++c)
            *entry(hi, hi, r, c) += AiT[3 * r] * Ji[c] + AiT[3 * r + 1] * Ji[3 + c] + AiT[3 * r + 2] * Ji[6 + c];
        }
      if (hj >= 0)
        for (int r = 0; r < 3; ++r) {
          b[3 * hj + r] -= AjT[3 * r] * err[0] + AjT[3 * r + 1] * err[1] + AjT[3 * r + 2] * err[2];
          for (int c = r; c < 3; ++c)
            *entry(hj, hj, r, c) += AjT[3 * r] * Jj[c] + AjT[3 * r + 1] * Jj[3 + c] + AjT[3 * r + 2] * Jj[6 + c];
        }
      if (hi >= 0 && hj >= 0 && hi != hj)
        // H_ij = Ji^T Omega Jj, stored in the upper block triangle (transposed when hi > hj)
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            const double v = AiT[3 * r] * Jj[c] + AiT[3 * r + 1] * Jj[3 + c] + AiT[3 * r + 2] * Jj[6 + c];
            if (hi < hj) *entry(hi, hj, r, c) += v;
            else *entry(hj, hi, c, r) += v;
          }
    }
    return chi2;
  }

  // n_iters Gauss-Newton iterations on `poses` (C7); returns the iterations done
  int iterate(double* poses, const double* meas, const double* info6, int n_iters, double* chi2_out) {
    std::vector<double> dx;
    int done = 0;
    for (int it = 0; it < n_iters; ++it) {
      double t0 = now();
      const double chi2 = linearise(poses, meas, info6);
      if (chi2_out) chi2_out[it] = chi2;
      t_lin += now() - t0;
      t0 = now();
      have_factor = n == 0 || numeric(Ax, &F);
      t_factor += now() - t0;
      if (!have_factor) break;
      t0 = now();
      if (n) solve(F, b, &dx);
      for (int v = 0; v < nv; ++v) {
        const int h = hidx[v];
        if (h < 0) continue;
        poses[3 * v] += dx[3 * h];  // C3
        poses[3 * v + 1] += dx[3 * h + 1];
        poses[3 * v + 2] = normalize_theta(poses[3 * v + 2] + dx[3 * h + 2]);
      }
      t_solve += now() - t0;
      ++done;
    }
    return done;
  }
};

}  // namespace

extern "C" {

// Gauss-Newton on an SE(2) pose graph. poses [nv][3] in/out; meas [ne][3]; info6 [ne][6] upper
// triangle row-major; fixed [nv] 0/1. chi2_out [n_iters] = chi2 at each iteration's linearisation
// point. times[0] = ordering + symbolic (once), [1] = linearise + build (sum), [2] = numeric
// factorisation (sum), [3] = triangular solves + update (sum). stats[0] = nnz(L), [1] = sum of
// squared column counts (~flops of one factorisation), [2] = 3 x free vertices.
// Returns iterations done (stops at the first failed factorisation, like optimize()).
int pgo_oracle_c_gauss_newton(int nv, int ne, const int32_t* ei, const int32_t* ej, const uint8_t* fixed,
                              double* poses, const double* meas, const double* info6, int n_iters,
                              double* chi2_out, double* times, long long* stats) {
  Graph g;
  g.set_structure(nv, ne, ei, ej, fixed);
  const int done = g.iterate(poses, meas, info6, n_iters, chi2_out);
  times[0] = g.t_analyse;
  times[1] = g.t_lin;
  times[2] = g.t_factor;
  times[3] = g.t_solve;
  stats[0] = g.F.Lp.empty() ? 0 : g.F.Lp[g.n];
  stats[1] = g.F.flops;
  stats[2] = g.n;
  return done;
}

}  // extern "C"

// ---- the pose-graph C ABI (include/pgo_solver.h) on the CPU -------------------------------------------
// So that the reference's own host sources, compiled against include/g2o_compat/g2o_compat.hpp, can
// be linked against THIS library instead of libcgmrslam_b200.so: the same front-end code then runs
// once over the CUDA solver and once over this oracle (oracle/Makefile, target `frontend`).
// Test infrastructure: the product never links it.
#include "../include/pgo_solver.h"

struct pgo_solver {
  Graph g;
  std::vector<double> poses, meas, info6;
  bool have_graph = false, have_values = false;
};

namespace {
thread_local std::string g_err;
int fail(int code, const char* msg) {
  g_err = msg;
  return code;
}
}  // namespace

extern "C" {

const char* pgo_last_error(void) { return g_err.c_str(); }

int pgo_create(pgo_solver** out, int, void*) {
  *out = new pgo_solver();
  return PGO_OK;
}
void pgo_destroy(pgo_solver* s) { delete s; }

int pgo_set_graph(pgo_solver* s, int nv, int ne, const int32_t* ei, const int32_t* ej, const uint8_t* fixed) {
  if (!s || nv <= 0 || ne < 0 || !fixed) return fail(PGO_ERR_ARG, "bad argument");
  for (int e = 0; e < ne; ++e)
    if (ei[e] < 0 || ei[e] >= nv || ej[e] < 0 || ej[e] >= nv || ei[e] == ej[e]) return fail(PGO_ERR_ARG, "bad edge");
  s->g.set_structure(nv, ne, ei, ej, fixed);
  s->have_graph = true;
  s->have_values = false;
  return PGO_OK;
}

int pgo_upload(pgo_solver* s, const double* poses, const double* meas, const double* info6) {
  if (!s || !s->have_graph) return fail(PGO_ERR_ARG, "pgo_set_graph first");
  s->poses.assign(poses, poses + 3 * static_cast<size_t>(s->g.nv));
  s->meas.assign(meas, meas + 3 * static_cast<size_t>(s->g.ne));
  s->info6.assign(info6, info6 + 6 * static_cast<size_t>(s->g.ne));
  s->have_values = true;
  s->g.have_factor = false;
  return PGO_OK;
}

int pgo_set_poses(pgo_solver* s, const double* poses) {
  if (!s || !s->have_values) return fail(PGO_ERR_ARG, "pgo_upload first");
  s->poses.assign(poses, poses + 3 * static_cast<size_t>(s->g.nv));
  return PGO_OK;
}

int pgo_get_poses(pgo_solver* s, double* poses) {
  if (!s || !s->have_values) return fail(PGO_ERR_ARG, "pgo_upload first");
  std::copy(s->poses.begin(), s->poses.end(), poses);
  return PGO_OK;
}

int pgo_iterate(pgo_solver* s, int n_iters, double* poses_out, double* chi2_out, int* iters_done) {
  if (!s || !s->have_values || n_iters < 0) return fail(PGO_ERR_ARG, "pgo_upload first");
  const int done = s->g.iterate(s->poses.data(), s->meas.data(), s->info6.data(), n_iters, chi2_out);
  if (iters_done) *iters_done = done;
  if (poses_out) std::copy(s->poses.begin(), s->poses.end(), poses_out);
  if (done < n_iters) return fail(PGO_ERR_NUMERIC, "H is not positive definite");
  return PGO_OK;
}

int pgo_chi2(pgo_solver* s, double* chi2) {
  if (!s || !s->have_values) return fail(PGO_ERR_ARG, "pgo_upload first");
  // linearise() overwrites the system of the last iteration: work on a copy of the numeric state
  Graph tmp = s->g;
  *chi2 = tmp.linearise(s->poses.data(), s->meas.data(), s->info6.data());
  return PGO_OK;
}

// C9: blocks (r, c) of H^-1 with the factor of the LAST iteration's H, unit right-hand sides
int pgo_marginals(pgo_solver* s, int n_blocks, const int32_t* vr, const int32_t* vc, double* cov_out) {
  if (!s || !s->have_values) return fail(PGO_ERR_ARG, "pgo_upload first");
  if (!s->g.have_factor) return fail(PGO_ERR_ARG, "no factorisation: call pgo_iterate first");
  const Graph& g = s->g;
  std::vector<double> rhs(g.n), x;
  for (int k = 0; k < n_blocks; ++k) {
    if (vr[k] < 0 || vr[k] >= g.nv || vc[k] < 0 || vc[k] >= g.nv || g.hidx[vr[k]] < 0 || g.hidx[vc[k]] < 0)
      return fail(PGO_ERR_ARG, "marginal requested for a fixed or unknown vertex");
    const int hr = g.hidx[vr[k]], hc = g.hidx[vc[k]];
    for (int c = 0; c < 3; ++c) {
      std::fill(rhs.begin(), rhs.end(), 0.0);
      rhs[3 * hc + c] = 1.0;
      solve(g.F, rhs, &x);
      for (int r = 0; r < 3; ++r) cov_out[9 * k + 3 * r + c] = x[3 * hr + r];
    }
  }
  return PGO_OK;
}

// C10: spanning tree from the fixed vertices, uniform edge cost. Ties independent of the order of the
// edge list: levels of a breadth-first search, a level's vertices in ascending index, a vertex's
// edges in ascending (other end, direction, measurement); the first edge to reach a vertex sets it.
int pgo_initial_guess(pgo_solver* s) {
  if (!s || !s->have_values) return fail(PGO_ERR_ARG, "pgo_upload first");
  const Graph& g = s->g;
  std::vector<std::vector<int> > adj(g.nv);
  for (int e = 0; e < g.ne; ++e) {
    adj[g.ei[e]].push_back(2 * e);      // this vertex is Xi
    adj[g.ej[e]].push_back(2 * e + 1);  // this vertex is Xj
  }
  const std::vector<double>& meas = s->meas;
  auto before = [&g, &meas](int a, int b) {
    const int ea = a >> 1, eb = b >> 1;
    const int wa = (a & 1) ? g.ei[ea] : g.ej[ea], wb = (b & 1) ? g.ei[eb] : g.ej[eb];
    if (wa != wb) return wa < wb;
    if ((a & 1) != (b & 1)) return (a & 1) < (b & 1);
    for (int k = 0; k < 3; ++k)
      if (meas[3 * ea + k] != meas[3 * eb + k]) return meas[3 * ea + k] < meas[3 * eb + k];
    return false;
  };
  for (int v = 0; v < g.nv; ++v) std::stable_sort(adj[v].begin(), adj[v].end(), before);
  std::vector<char> seen(g.nv, 0);
  std::vector<int> level, next;
  for (int v = 0; v < g.nv; ++v)
    if (g.fixed[v]) {
      seen[v] = 1;
      level.push_back(v);
    }
  while (!level.empty()) {
    next.clear();
    for (size_t q = 0; q < level.size(); ++q) {
      const int v = level[q];
      const SE2 xv = {s->poses[3 * v], s->poses[3 * v + 1], s->poses[3 * v + 2]};
      for (size_t t = 0; t < adj[v].size(); ++t) {
        const int e = adj[v][t] >> 1;
        const bool forward = !(adj[v][t] & 1);
        const int w = forward ? g.ej[e] : g.ei[e];
        if (seen[w]) continue;
        const SE2 z = {s->meas[3 * e], s->meas[3 * e + 1], s->meas[3 * e + 2]};
        const SE2 xw = forward ? mul(xv, z) : mul(xv, inv(z));  // EdgeSE2::initialEstimate (C4)
        s->poses[3 * w] = xw.x;
        s->poses[3 * w + 1] = xw.y;
        s->poses[3 * w + 2] = xw.th;
        seen[w] = 1;
        next.push_back(w);
      }
    }
    std::sort(next.begin(), next.end());
    level.swap(next);
  }
  return PGO_OK;
}

// C11: EdgeLabeler::labelEdges for star edges gauge -> v with the gauge fixed
int pgo_label_star_edges(pgo_solver* s, int gauge, int n, const int32_t* v, double* meas_out, double* info_out) {
  if (!s || !s->have_values || gauge < 0 || gauge >= s->g.nv) return fail(PGO_ERR_ARG, "bad argument");
  if (n == 0) return PGO_OK;
  std::vector<double> cov(9 * static_cast<size_t>(n));
  const int rc = pgo_marginals(s, n, v, v, cov.data());
  if (rc) return rc;
  const double alpha = 1e-3, beta = 2.0, dim = 3.0, lam = alpha * alpha * dim;
  const double wi = 1.0 / (2.0 * (dim + lam)), w0m = lam / (dim + lam), w0c = w0m + (1.0 - alpha * alpha + beta);
  const SE2 xg = {s->poses[3 * gauge], s->poses[3 * gauge + 1], s->poses[3 * gauge + 2]};
  const SE2 xgi = inv(xg);
  for (int t = 0; t < n; ++t) {
    const SE2 xv = {s->poses[3 * v[t]], s->poses[3 * v[t] + 1], s->poses[3 * v[t] + 2]};
    const SE2 z = mul(xgi, xv);  // setMeasurementFromState
    const SE2 zi = inv(z);
    double a[9], L[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 9; ++i) a[i] = cov[9 * t + i] * (dim + lam);
    if (!(a[0] > 0.0)) return fail(PGO_ERR_NUMERIC, "covariance not positive definite");
    L[0] = std::sqrt(a[0]);
    L[3] = a[3] / L[0];
    L[6] = a[6] / L[0];
    const double d1 = a[4] - L[3] * L[3];
    if (!(d1 > 0.0)) return fail(PGO_ERR_NUMERIC, "covariance not positive definite");
    L[4] = std::sqrt(d1);
    L[7] = (a[7] - L[6] * L[3]) / L[4];
    const double d2 = a[8] - L[6] * L[6] - L[7] * L[7];
    if (!(d2 > 0.0)) return fail(PGO_ERR_NUMERIC, "covariance not positive definite");
    L[8] = std::sqrt(d2);
    double errs[7][3], mean[3] = {0, 0, 0}, S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 7; ++k) {
      SE2 pv = xv;
      if (k > 0) {
        const int col = (k - 1) / 2;
        const double sgn = (k - 1) % 2 ? -1.0 : 1.0;
        pv.x += sgn * L[col];
        pv.y += sgn * L[3 + col];
        pv.th = normalize_theta(pv.th + sgn * L[6 + col]);
      }
      const SE2 e = mul(zi, mul(xgi, pv));
      errs[k][0] = e.x;
      errs[k][1] = e.y;
      errs[k][2] = e.th;
      for (int i = 0; i < 3; ++i) mean[i] += (k ? wi : w0m) * errs[k][i];
    }
    for (int k = 0; k < 7; ++k)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) S[3 * i + j] += (k ? wi : w0c) * (errs[k][i] - mean[i]) * (errs[k][j] - mean[j]);
    // information = S^-1 (3x3 symmetric, by cofactors)
    const double c00 = S[4] * S[8] - S[5] * S[7], c01 = S[5] * S[6] - S[3] * S[8], c02 = S[3] * S[7] - S[4] * S[6];
    const double det = S[0] * c00 + S[1] * c01 + S[2] * c02;
    if (!(det > 0.0)) return fail(PGO_ERR_NUMERIC, "sample covariance not positive definite");
    double* o = info_out + 9 * static_cast<size_t>(t);
    o[0] = c00 / det;
    o[1] = o[3] = c01 / det;
    o[2] = o[6] = c02 / det;
    o[4] = (S[0] * S[8] - S[2] * S[6]) / det;
    o[5] = o[7] = (S[2] * S[3] - S[0] * S[5]) / det;
    o[8] = (S[0] * S[4] - S[1] * S[3]) / det;
    meas_out[3 * t] = z.x;
    meas_out[3 * t + 1] = z.y;
    meas_out[3 * t + 2] = z.th;
  }
  return PGO_OK;
}

}  // extern "C"
