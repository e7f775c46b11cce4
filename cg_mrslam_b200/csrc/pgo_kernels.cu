// CUDA back-end of the SE(2) pose-graph solver (sm_100a). Implements pgo_device.h.
//
// A Gauss-Newton iteration -- the work of SparseOptimizer::optimize(1) with BlockSolver +
// LinearSolverCSparse + OptimizationAlgorithmGaussNewton (configured by
// the reference at src/slam/graph_slam.cpp:44-56) -- is one CUDA graph of per-phase kernels:
//   1. gn_linearise: per free vertex, every incident EdgeSE2 is evaluated (error e, analytic
//      Jacobians Ji, Jj, C4) and H_pp += Jp^T Om Jp, b_p += Jp^T (-Om e), H_qp += Jq^T Om Jp are
//      accumulated by the one thread that owns the destination block: no atomics, fixed order
//      (C5). chi2 = sum e^T Om e is reduced deterministically.
//   2. factorise + forward-substitute, bottom-up over the panel levels: supernodal block L D L^T of
//      the permuted H (pgo_supernodal.h): sn_k_factor / sn_k_fused per level, then sn_k_update, the
//      outer products on the fp64 tensor cores.
//   3. back-substitute, top-down over the supernode levels: sn_k_bwd_*.
//   4. gn_update: VertexSE2::oplusImpl (C3) on every free vertex.
// Marginal covariance blocks are extra right-hand sides against the stored factor (stand-alone
// forward + backward substitution with the same task functions); pgo_dd.cuh holds what the
// domain-decomposed multi-GPU iteration adds (owned-edge linearisation, exchange packing).
// Algorithmic bytes and measured shares per kernel are stated in DESIGN.md section 4.3.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/pgo_solver.h"
#include "pgo_device.h"
#include "pgo_supernodal.h"

namespace cg = cooperative_groups;

namespace pgo {

namespace {

#define PGO_CUDA(call)                                                   \
  do {                                                                   \
    cudaError_t e_ = (call);                                             \
    if (e_ != cudaSuccess) {                                             \
      if (err) *err = std::string(#call) + ": " + cudaGetErrorString(e_); \
      return PGO_ERR_CUDA;                                               \
    }                                                                    \
  } while (0)

const double kPi = 3.14159265358979323846;

struct Params {
  // graph
  int n_vertices, n_edges, n;  // n = free vertices
  const int* edge_i;
  const int* edge_j;
  const int* vpos;
  const int* inc_ptr;
  const Incidence* inc;
  const int* ff_edges;
  int n_ff;
  // warp-cooperative linearisation: incidence entries regrouped per warp (see gn_linearise)
  const int4* winc;      // {edge, pos, p, role | count_chi2 << 1}
  const int2* winc_ends; // {edge_i, edge_j} of the entry's edge
  const int* wg_ptr;     // [n_wgroups + 1] entry ranges; a range of more than 32 entries is ONE vertex
  const int* wg_info;    // bit 0: some vertex of the group has two entries with one pos (parallel edges)
  int n_wgroups;
  double* poses;
  const double* meas;
  const double* info6;
  // factor structure
  long long nnzb;
  const int* col_ptr;
  const int* row_idx;
  const int* perm_vertex;  // permuted position -> vertex index
  // numeric
  double* M;      // [nnzb][9]
  double* Dinv;   // [n][9]
  double* rhs;    // [n][3]
  double* u;      // [nrhs][n][3]
  double* x;      // [nrhs][n][3]
  double* chi2_partial;  // [gridDim.x]
  double* trig;          // [n_vertices + n_edges][2]: sin, cos of every pose angle / measurement angle
  double* chi2_out;      // [n_iters]
  int* status;           // [0] error flag, [1] iterations done
  unsigned long long* stamps;  // [6] globaltimer at the stage boundaries of the last iteration
  // batch of problem instances with this structure: element strides between instances
  long long s_poses, s_meas, s_info, s_M, s_Dinv, s_vec, s_partial, s_chi2, s_trig;
};

// Instance b of a batch (blockIdx.y in the kernels of the iteration graph).
__device__ __forceinline__ Params params_at(Params P, int b) {
  P.poses += b * P.s_poses;
  P.meas += b * P.s_meas;
  P.info6 += b * P.s_info;
  P.M += b * P.s_M;
  P.Dinv += b * P.s_Dinv;
  P.rhs += b * P.s_vec;
  P.u += b * P.s_vec;
  P.x += b * P.s_vec;
  P.chi2_partial += b * P.s_partial;
  P.trig += b * P.s_trig;
  P.chi2_out += b * P.s_chi2;
  P.status += 4 * b;
  return P;
}

// ---- small dense helpers, 3x3 row-major ----------------------------------------------------------
// Inverse of a symmetric positive definite 3x3 via Cholesky; false if a pivot is not positive.
__device__ bool spd_inverse(const double* a, double* inv) {
  const double l00sq = a[0];
  if (!(l00sq > 0.0)) return false;
  const double l00 = sqrt(l00sq);
  const double l10 = a[3] / l00, l20 = a[6] / l00;
  const double l11sq = a[4] - l10 * l10;
  if (!(l11sq > 0.0)) return false;
  const double l11 = sqrt(l11sq);
  const double l21 = (a[7] - l20 * l10) / l11;
  const double l22sq = a[8] - l20 * l20 - l21 * l21;
  if (!(l22sq > 0.0)) return false;
  const double l22 = sqrt(l22sq);
  // inverse of the lower-triangular factor
  const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
  const double i10 = -l10 * i00 * i11;
  const double i21 = -l21 * i11 * i22;
  const double i20 = -(l20 * i00 + l21 * i10) * i22;
  // A^-1 = Linv^T Linv
  inv[0] = i00 * i00 + i10 * i10 + i20 * i20;
  inv[1] = inv[3] = i10 * i11 + i20 * i21;
  inv[2] = inv[6] = i20 * i22;
  inv[4] = i11 * i11 + i21 * i21;
  inv[5] = inv[7] = i21 * i22;
  inv[8] = i22 * i22;
  return true;
}

// g2o normalize_theta: identity inside [-pi, pi), else wrap.
__device__ __forceinline__ double normalize_theta(double t) {
  if (t >= -kPi && t < kPi) return t;
  const double two_pi = 2.0 * kPi;
  double w = t - two_pi * floor(t / two_pi);
  if (w >= kPi) w -= two_pi;
  return w;
}

struct SE2d {
  double x, y, th;
};

__device__ __forceinline__ SE2d se2_inv(const SE2d& a) {  // C1
  SE2d r;
  r.th = normalize_theta(-a.th);
  double s, c;
  sincos(r.th, &s, &c);
  r.x = c * (-a.x) - s * (-a.y);
  r.y = s * (-a.x) + c * (-a.y);
  return r;
}

__device__ __forceinline__ SE2d se2_mul(const SE2d& a, const SE2d& b) {  // C1
  double s, c;
  sincos(a.th, &s, &c);
  SE2d r;
  r.x = a.x + (c * b.x - s * b.y);
  r.y = a.y + (s * b.x + c * b.y);
  r.th = normalize_theta(a.th + b.th);
  return r;
}

__device__ __forceinline__ SE2d load_pose(const double* p, int v) {
  SE2d r = {p[3 * v], p[3 * v + 1], p[3 * v + 2]};
  return r;
}

// EdgeSE2::computeError + linearizeOplus (C4) + Omega of one edge.
struct EdgeLin {
  double e[3];
  double ji[9], jj[9];
  double om[9];
};

__device__ void linearise_edge(const Params& P, int edge, EdgeLin* L) {
  const SE2d xi = load_pose(P.poses, P.edge_i[edge]);
  const SE2d xj = load_pose(P.poses, P.edge_j[edge]);
  const SE2d z = {P.meas[3 * edge], P.meas[3 * edge + 1], P.meas[3 * edge + 2]};
  const SE2d zinv = se2_inv(z);
  const SE2d d = se2_mul(se2_inv(xi), xj);
  const SE2d er = se2_mul(zinv, d);
  L->e[0] = er.x;
  L->e[1] = er.y;
  L->e[2] = er.th;
  double si, ci;
  sincos(xi.th, &si, &ci);
  const double dx = xj.x - xi.x, dy = xj.y - xi.y;
  const double a[9] = {-ci, -si, -si * dx + ci * dy, si, -ci, -ci * dx - si * dy, 0.0, 0.0, -1.0};
  const double b[9] = {ci, si, 0.0, -si, ci, 0.0, 0.0, 0.0, 1.0};
  double sz, cz;
  sincos(zinv.th, &sz, &cz);
  // Rz = diag(R(theta(Z^-1)), 1); J <- Rz * J
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    L->ji[c] = cz * a[c] - sz * a[3 + c];
    L->ji[3 + c] = sz * a[c] + cz * a[3 + c];
    L->ji[6 + c] = a[6 + c];
    L->jj[c] = cz * b[c] - sz * b[3 + c];
    L->jj[3 + c] = sz * b[c] + cz * b[3 + c];
    L->jj[6 + c] = b[6 + c];
  }
  const double* w = P.info6 + 6 * static_cast<size_t>(edge);
  L->om[0] = w[0];
  L->om[1] = L->om[3] = w[1];
  L->om[2] = L->om[6] = w[2];
  L->om[4] = w[3];
  L->om[5] = L->om[7] = w[4];
  L->om[8] = w[5];
}

// The same with the sines / cosines taken from P.trig (gn_trig runs first): an iteration evaluates
// every edge from both of its ends, and each evaluation needs the rotations of theta_i, -theta_i,
// theta_z and -theta_z; with the table an iteration costs V + E sincos instead of ~10 E.
__device__ void linearise_edge_tab(const Params& P, int edge, int vi, int vj, EdgeLin* L) {
  const SE2d xi = load_pose(P.poses, vi);
  const SE2d xj = load_pose(P.poses, vj);
  const double si = P.trig[2 * vi], ci = P.trig[2 * vi + 1];
  const double* tz = P.trig + 2 * (static_cast<size_t>(P.n_vertices) + edge);
  const double sz = -tz[0], cz = tz[1];  // rotation of Z^-1
  const SE2d z = {P.meas[3 * edge], P.meas[3 * edge + 1], P.meas[3 * edge + 2]};
  SE2d zinv, ixi, d;
  zinv.th = normalize_theta(-z.th);
  zinv.x = cz * (-z.x) - sz * (-z.y);
  zinv.y = sz * (-z.x) + cz * (-z.y);
  {
    const double s = -si, c = ci;  // rotation of Xi^-1
    ixi.th = normalize_theta(-xi.th);
    ixi.x = c * (-xi.x) - s * (-xi.y);
    ixi.y = s * (-xi.x) + c * (-xi.y);
    d.x = ixi.x + (c * xj.x - s * xj.y);
    d.y = ixi.y + (s * xj.x + c * xj.y);
    d.th = normalize_theta(ixi.th + xj.th);
  }
  L->e[0] = zinv.x + (cz * d.x - sz * d.y);
  L->e[1] = zinv.y + (sz * d.x + cz * d.y);
  L->e[2] = normalize_theta(zinv.th + d.th);
  const double dx = xj.x - xi.x, dy = xj.y - xi.y;
  const double a[9] = {-ci, -si, -si * dx + ci * dy, si, -ci, -ci * dx - si * dy, 0.0, 0.0, -1.0};
  const double b[9] = {ci, si, 0.0, -si, ci, 0.0, 0.0, 0.0, 1.0};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    L->ji[c] = cz * a[c] - sz * a[3 + c];
    L->ji[3 + c] = sz * a[c] + cz * a[3 + c];
    L->ji[6 + c] = a[6 + c];
    L->jj[c] = cz * b[c] - sz * b[3 + c];
    L->jj[3 + c] = sz * b[c] + cz * b[3 + c];
    L->jj[6 + c] = b[6 + c];
  }
  const double* w = P.info6 + 6 * static_cast<size_t>(edge);
  L->om[0] = w[0];
  L->om[1] = L->om[3] = w[1];
  L->om[2] = L->om[6] = w[2];
  L->om[4] = w[3];
  L->om[5] = L->om[7] = w[4];
  L->om[8] = w[5];
}

__device__ __forceinline__ double edge_chi2(const EdgeLin& L) {
  double c = 0.0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) t += L.om[3 * r + k] * L.e[k];
    c += L.e[r] * t;
  }
  return c;
}

// Deterministic block reduction of one double per thread; result valid in thread 0.
__device__ double block_sum(double v, double* scratch) {
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) r += scratch[w];
  __syncthreads();
  return r;
}

// ---- phase 1: linearise (gn_linearise below) -----------------------------------------------------
const int kThreads = 256;
// linearise: one warp per group of vertices (<= 32 incidences), 4 warps per CTA
const int kLinThreads = 128;

// ---- supernodal single-GPU iteration (pgo_supernodal.h) ------------------------------------------
// One Gauss-Newton iteration is a CUDA graph of small kernels, one per phase and panel / supernode
// level: a kernel boundary is the barrier between dependent phases, every task is one CTA (or one
// warp of a 4-warp CTA) so the hardware block scheduler balances the load, each kernel gets its
// own register and shared-memory budget, and ncu sees every phase as a launch of its own.
struct CtaGroup {
  __device__ __forceinline__ int rank() const { return threadIdx.x; }
  __device__ __forceinline__ int size() const { return blockDim.x; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};
struct WarpGroup {
  __device__ __forceinline__ int rank() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int size() const { return 32; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};

const int kCtaThreads = 256;   // CTA tasks
const int kUpdateThreads = 128;  // outer-product tiles: 4 warps, two 8-row strips each
const int kWarpsPerCta = 4;    // warp tasks: 4 per CTA

__device__ __forceinline__ bool sn_failed(const SNView& V) {
  return *reinterpret_cast<volatile int*>(V.status) != 0;
}

// fa: panel factorisation, one CTA per (panel, row chunk)
__global__ void __launch_bounds__(kCtaThreads) sn_k_factor(SNView V, const Task* tasks) {
  extern __shared__ double sm[];
  V = sn_at_instance(V, blockIdx.y);
  if (sn_failed(V)) return;
  sn_task_factor(CtaGroup(), V, tasks[blockIdx.x], sm);
}
// fb: outer-product tiles, one CTA per tile, on the fp64 tensor cores.
//   C[3 ti x 3 tj] = A[3 ti x 3 w] * B[3 w x 3 tj],  A = the scaled blocks Y(a,t) = M(a,t) Dinv_t of
//   the panel's rows a, B = the blocks M(b,t)^T of the rows b;  M(r_a, r_b) -= C(a,b) (fp64 atomics).
// This is the one true GEMM of the solver (north star: "tensor cores only where it is a true GEMM").
// mma.sync.m8n8k4.f64 reaches the same 37 TFLOP/s as DFMA on this part (tools/ubench/fp64_rate.cu)
// with an eighth of the instructions and one shared load per 8x4 fragment instead of two per
// multiply-add, which is what the scalar version of this kernel was bound by (ncu: fp64 pipe 23 %).
// Shared memory: As[3 ti][ldk], Bt[3 tj][ldk] (B transposed: both fragments read row-major with
// the k index fastest; ldk = 4 mod 16 doubles keeps a fragment's 32 loads on distinct banks),
// tp[ti * tj] scatter positions (-1 = no target).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}

// One tile of an outer product: rows [i0, i0 + ti) x columns [j0, j0 + tj) of the below-row list
// of the LAST panel of the task's panel range, summed over the columns of all panels of the range
// (one panel for the updates inside a supernode, all of a supernode's panels for the update of its
// ancestors), one 16-column chunk staged at a time, accumulators in registers across the chunks.
template <int NT>  // column tiles of 8 scalars: tj = 8 NT / 3 blocks
__device__ __forceinline__ void update_tile(const SNView& V, const Task& T, double* sm) {
  const int ti = T.aux & 0xFF, tj = (T.aux >> 8) & 0xFF, np = (T.aux >> 16) & 0x3F;
  const int q_last = T.id - ((T.aux >> 22) & 0x3F), q_first = q_last - np + 1;
  const bool inside = (T.aux >> 28) & 1;  // targets = later columns of the same supernode only
  const int i0 = T.r0, j0 = T.r1;
  const PanelDesc last = V.pn[T.id];
  const int m = last.m;
  // inside a supernode the columns b stop at the supernode's last column
  const SuperDesc sd = V.sn[last.sn];
  const int jlim = inside ? sd.W - last.sn_off - last.w : m;
  const int ni = min(m - i0, ti), nj = min(jlim - j0, tj);
  const int ldk = sn_tile_ld(!inside && sd.pn_end - sd.pn_begin > 1 ? kPanelWidth : last.w);
  double* As = sm;
  double* Bt = As + 3 * ti * ldk;
  int* tp = reinterpret_cast<int*>(Bt + 3 * tj * ldk);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lx = tid & 31, wy = tid >> 5, ny = nthr >> 5;
  const int lane = lx, warp = wy, fr = lane >> 2, fk = lane & 3;
  const int mt = (3 * ni + 7) / 8;
  if (i0 == 0 && j0 == 0 && q_last == T.id && last.scratch >= 0) {  // move a scratch-published diagonal part into place
    const int w = last.w, len = w + m;
    const double* src = V.scratch + 9 * static_cast<size_t>(last.scratch);
    for (int idx = tid; idx < w * w * 9; idx += nthr) {
      const int i = idx / (w * 9), t = (idx / 9) % w, k = idx % 9;
      if (i >= t) V.M[9 * static_cast<size_t>(sn_colpos(last.base, len, t) + (i - t)) + k] = __ldcg(src + idx);
    }
  }
  // scatter positions of the tile's block pairs: the column's two look-ups once per lane, then one
  // independent look-up per row (all staging loops are two-dimensional: no integer division)
  if (lx < tj) {
    const int b = j0 + lx;
    int cb = 0, to = 0;
    if (lx < nj) {
      cb = V.colbase[last.meta + b];
      to = V.tbl_off[last.meta + b];
    }
    for (int al = wy; al < ti; al += ny) {
      const int a = i0 + al;
      int pos = -1;
      if (al < ni && lx < nj && a >= b) pos = cb + V.tbl[to + a];
      tp[al * tj + lx] = pos;
    }
  }
  double acc[2][NT][2];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[h][n][0] = acc[h][n][1] = 0.0;
  for (int q = q_first; q <= q_last; ++q) {
    const PanelDesc pd = V.pn[q];
    const int w = pd.w, len = w + pd.m;
    const int off = pd.m - m;  // the last panel's rows start here in this panel's below list
    const int k4 = (3 * w + 3) & ~3;
    if (q > q_first) __syncthreads();  // everybody is done with the previous chunk
    // zero the k padding (3 w .. k4) of both operands
    for (int c = 3 * w + wy; c < k4; c += ny)
      for (int row = lx; row < 3 * (ti + tj); row += 32) {
        if (row < 3 * ti) As[row * ldk + c] = 0.0;
        else Bt[(row - 3 * ti) * ldk + c] = 0.0;
      }
    // A = Y rows (scaled by the panel factorisation), B transposed = M rows, copied into row-major
    // [3 block + row][3 t + column] with asynchronous 8-byte global->shared copies so that the whole
    // chunk is in flight at once. One warp per (operand, column): the tile's blocks of a column are
    // ONE contiguous run of 9 n doubles in global memory, and the lanes walk it double by double, so
    // a warp instruction reads 256 consecutive bytes = 8 full sectors (the first version had the
    // lanes on different blocks: 32 sectors per instruction, a quarter of each used, and the L1
    // data path 81 % busy -- profiles/r02_ncu_sn_k_update_b128.txt). Double q of the run is row
    // q / 3, column q % 3 of the operand: destination (q / 3) ldk + 3 t + q % 3.
    for (int c = wy; c < 2 * w; c += ny) {
      const bool is_a = c < w;
      const int t = is_a ? c : c - w;
      const int n9 = 9 * (is_a ? ni : nj);
      const size_t col = static_cast<size_t>(sn_colpos(pd.base, len, t) + (w - t) + off);
      const double* src = (is_a ? V.Y + 9 * static_cast<size_t>(i0) : V.M + 9 * static_cast<size_t>(j0)) + 9 * col;
      double* dst = (is_a ? As : Bt) + 3 * t;
      for (int q = lx; q < n9; q += 32) {
        const int row = q / 3;
        cp_async8(dst + row * ldk + (q - 3 * row), src + q);
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // a warp's two 8-row strips share every B fragment: 2 + NT shared-memory loads per k step for
    // 2 NT DMMAs (the loop is bound by shared-memory bandwidth: ncu, 0.74 wavefronts / clock / SM)
    {
      const double* bp = Bt + fr * ldk + fk;
      const double* ap0 = As + (8 * warp + fr) * ldk + fk;
      const double* ap1 = As + (8 * (warp + ny) + fr) * ldk + fk;
      if (warp + ny < mt) {
        for (int k0 = 0; k0 < k4; k0 += 4) {
          const double a0 = ap0[k0], a1 = ap1[k0];
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const double b = bp[8 * n * ldk + k0];
            dmma884(acc[0][n][0], acc[0][n][1], a0, b);
            dmma884(acc[1][n][0], acc[1][n][1], a1, b);
          }
        }
      } else if (warp < mt) {
        for (int k0 = 0; k0 < k4; k0 += 4) {
          const double a0 = ap0[k0];
#pragma unroll
          for (int n = 0; n < NT; ++n) dmma884(acc[0][n][0], acc[0][n][1], a0, bp[8 * n * ldk + k0]);
        }
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int mi = warp + ny * h;
    if (mi >= mt) continue;
    const int row = 8 * mi + fr, al = row / 3, r = row - 3 * al;
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = 8 * n + 2 * fk + e, bl = col / 3, c = col - 3 * bl;
        const int pos = tp[al * tj + bl];
        if (pos >= 0) atomicAdd(V.M + 9 * static_cast<size_t>(pos) + 3 * r + c, -acc[h][n][e]);
      }
  }
}

__global__ void __launch_bounds__(kUpdateThreads, 5) sn_k_update(SNView V, const Task* tasks) {
  extern __shared__ double sm[];
  V = sn_at_instance(V, blockIdx.y);
  if (sn_failed(V)) return;
  const Task T = tasks[blockIdx.x];
  switch ((3 * ((T.aux >> 8) & 0xFF)) / 8) {
    case 3: update_tile<3>(V, T, sm); break;
    case 6: update_tile<6>(V, T, sm); break;
    default: update_tile<9>(V, T, sm); break;
  }
}
// ff: small panels start to finish, one warp each
// (stride = shared-memory doubles per warp: the largest task of the level)
__global__ void __launch_bounds__(32 * kWarpsPerCta, 8) sn_k_fused(SNView V, const Task* tasks, int n, int stride) {
  extern __shared__ double sm[];
  const int i = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  V = sn_at_instance(V, blockIdx.y);
  if (i >= n || sn_failed(V)) return;
  sn_task_fused(WarpGroup(), V, tasks[i], sm + (threadIdx.x >> 5) * stride);
}
// backward substitution
__global__ void __launch_bounds__(32 * kWarpsPerCta) sn_k_bwd_rows(SNView V, const Task* tasks, int n) {
  const int i = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  V = sn_at_instance(V, blockIdx.y);
  if (i >= n || sn_failed(V)) return;
  sn_backward_rows(WarpGroup(), V, tasks[i].id, tasks[i].r0, tasks[i].r1);
}
__global__ void __launch_bounds__(32 * kWarpsPerCta, 8) sn_k_bwd_small(SNView V, const Task* tasks, int n, int stride) {
  extern __shared__ double sm[];
  const int i = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  V = sn_at_instance(V, blockIdx.y);
  if (i >= n || sn_failed(V)) return;
  sn_task_backward_small(WarpGroup(), V, tasks[i], sm + (threadIdx.x >> 5) * stride);
}
__global__ void __launch_bounds__(kCtaThreads, 2) sn_k_bwd_tri(SNView V, const Task* tasks) {
  extern __shared__ double sm[];
  V = sn_at_instance(V, blockIdx.y);
  if (sn_failed(V)) return;
  sn_task_backward_tri(CtaGroup(), V, tasks[blockIdx.x], sm);
}

// stand-alone forward substitution (right-hand sides solved after the factorisation: marginals)
__global__ void __launch_bounds__(kCtaThreads) sn_k_fwd_tri(SNView V, const Task* tasks) {
  extern __shared__ double sm[];
  V = sn_at_instance(V, blockIdx.y);
  sn_task_forward_tri(CtaGroup(), V, tasks[blockIdx.x], sm);
}
__global__ void __launch_bounds__(32 * kWarpsPerCta) sn_k_fwd_small(SNView V, const Task* tasks, int n, int stride) {
  extern __shared__ double sm[];
  const int i = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  V = sn_at_instance(V, blockIdx.y);
  if (i >= n) return;
  sn_task_forward_small(WarpGroup(), V, tasks[i], sm + (threadIdx.x >> 5) * stride);
}
__global__ void __launch_bounds__(32 * kWarpsPerCta) sn_k_fwd_rows(SNView V, const Task* tasks, int n) {
  const int i = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  V = sn_at_instance(V, blockIdx.y);
  if (i >= n) return;
  sn_forward_rows(WarpGroup(), V, tasks[i].id, tasks[i].r0, tasks[i].r1, nullptr);
}

// linearise + assemble (phase 1): g2o's EdgeSE2::linearizeOplus + constructQuadraticForm + buildSystem
//  as ONE warp-cooperative kernel. Every incidence (edge, end) is one lane: the lane
// evaluates the edge (error, both Jacobians: sines / cosines come from the gn_trig table, edge data
// are 72-byte gathers) and forms its contribution to the Hessian row of its own end -- the diagonal
// block, the right-hand side, and the off-diagonal block this end owns. Lanes are laid out so that
// the incidences of a vertex are consecutive (host: dev_set_structure), and the contributions are
// summed per vertex with a segmented warp-shuffle reduction (5 shuffle steps, fixed order: bitwise
// reproducible); only the head lane of a vertex writes H_pp and b_p. Off-diagonal blocks have one
// contributing lane each and are written directly -- except parallel edges (two constraints between
// one pair of vertices), whose lanes are adjacent and go through the same reduction keyed by the
// block position. No atomics. A vertex of degree > 32 is a group of its own, reduced round by round.
__device__ __forceinline__ double shfl_down_d(double v, int off) { return __shfl_down_sync(0xffffffffu, v, off); }

__global__ void __launch_bounds__(kLinThreads, 5) gn_linearise(Params P) {
  P = params_at(P, blockIdx.y);
  if (*reinterpret_cast<volatile int*>(P.status) != 0) return;
  const int lane = threadIdx.x & 31, g = blockIdx.x * (kLinThreads / 32) + (threadIdx.x >> 5);
  double chi = 0.0;
  if (g < P.n_wgroups) {
    const int e0 = P.wg_ptr[g], e1 = P.wg_ptr[g + 1];
    const bool dup = P.wg_info[g] & 1, big = e1 - e0 > 32;
    double carry[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // lane 0 of a big vertex: h00 h10 h11 h20 h21 h22 b0 b1 b2
    int p_big = -1;
    for (int base = e0; base < e1; base += 32) {
      const int idx = base + lane;
      const bool act = idx < e1;
      int4 en = make_int4(0, -1, -2 - lane, 0);
      if (act) en = P.winc[idx];
      double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, off[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      if (act) {
        EdgeLin L;
        const int2 ends = P.winc_ends[idx];  // the edge's two vertices: no dependent look-up through edge_i / edge_j
        linearise_edge_tab(P, en.x, ends.x, ends.y, &L);
        const bool role = en.w & 1;
        const double* jo = role ? L.jj : L.ji;  // this vertex
        const double* jx = role ? L.ji : L.jj;  // the other one
        double A[9];                            // Jo^T Omega
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            A[3 * r + c] = jo[r] * L.om[c] + jo[3 + r] * L.om[3 + c] + jo[6 + r] * L.om[6 + c];
        int k = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c <= r; ++c)  // lower triangle of Jo^T Omega Jo
            v[k++] = A[3 * r] * jo[c] + A[3 * r + 1] * jo[3 + c] + A[3 * r + 2] * jo[6 + c];
#pragma unroll
        for (int r = 0; r < 3; ++r) v[6 + r] = -(A[3 * r] * L.e[0] + A[3 * r + 1] * L.e[1] + A[3 * r + 2] * L.e[2]);
        if (en.y >= 0)  // block (row = other vertex, column = this vertex) = Jx^T Omega Jo = (A Jx)^T
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
              off[3 * c + r] = A[3 * r] * jx[c] + A[3 * r + 1] * jx[3 + c] + A[3 * r + 2] * jx[6 + c];
        if (en.w & 2) chi += edge_chi2(L);
      }
      // per-vertex sums: segmented reduction over the lanes with the same p
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int pk = __shfl_down_sync(0xffffffffu, en.z, o);
        const bool take = lane + o < 32 && pk == en.z;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const double other = shfl_down_d(v[i], o);
          if (take) v[i] += other;
        }
      }
      if (dup) {  // parallel edges: the same with the block position as the key
        int key = en.y >= 0 ? en.y : -2 - lane;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int pk = __shfl_down_sync(0xffffffffu, key, o);
          const bool take = lane + o < 32 && pk == key;
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const double other = shfl_down_d(off[i], o);
            if (take) off[i] += other;
          }
        }
        const int prev = __shfl_up_sync(0xffffffffu, key, 1);
        if (lane > 0 && prev == key) en.y = -1;  // not the head of its run: nothing to write
      }
      if (act && en.y >= 0) {
        double* dst = P.M + 9 * static_cast<size_t>(en.y);
#pragma unroll
        for (int i = 0; i < 9; ++i) dst[i] = off[i];
      }
      const int prev_p = __shfl_up_sync(0xffffffffu, en.z, 1);
      const bool head = act && (lane == 0 || prev_p != en.z);
      if (big) {
        if (lane == 0) {
          p_big = en.z;
#pragma unroll
          for (int i = 0; i < 9; ++i) carry[i] += v[i];
        }
      } else if (head) {
        double* d = P.M + 9 * static_cast<size_t>(P.col_ptr[en.z]);
        d[0] = v[0];
        d[1] = d[3] = v[1];
        d[4] = v[2];
        d[2] = d[6] = v[3];
        d[5] = d[7] = v[4];
        d[8] = v[5];
        P.rhs[3 * en.z] = v[6];
        P.rhs[3 * en.z + 1] = v[7];
        P.rhs[3 * en.z + 2] = v[8];
      }
    }
    if (big && lane == 0) {
      double* d = P.M + 9 * static_cast<size_t>(P.col_ptr[p_big]);
      d[0] = carry[0];
      d[1] = d[3] = carry[1];
      d[4] = carry[2];
      d[2] = d[6] = carry[3];
      d[5] = d[7] = carry[4];
      d[8] = carry[5];
      P.rhs[3 * p_big] = carry[6];
      P.rhs[3 * p_big + 1] = carry[7];
      P.rhs[3 * p_big + 2] = carry[8];
    }
    // edges between two fixed vertices only add to chi2: dealt to the first groups, one per lane
    for (int t = g * 32 + lane; t < P.n_ff; t += P.n_wgroups * 32) {
      EdgeLin L;
      const int e = P.ff_edges[t];
      linearise_edge_tab(P, e, P.edge_i[e], P.edge_j[e], &L);
      chi += edge_chi2(L);
    }
  }
  for (int o = 16; o; o >>= 1) chi += __shfl_down_sync(0xffffffffu, chi, o);
  if (lane == 0) P.chi2_partial[blockIdx.x * (kLinThreads / 32) + (threadIdx.x >> 5)] = chi;
}
// sin / cos of every pose angle and every measurement angle of this iteration (see linearise_edge_tab)
__global__ void gn_trig(Params P) {
  P = params_at(P, blockIdx.y);
  if (*reinterpret_cast<volatile int*>(P.status) != 0) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.n_vertices + P.n_edges) return;
  const double th = t < P.n_vertices ? P.poses[3 * static_cast<size_t>(t) + 2]
                                     : P.meas[3 * static_cast<size_t>(t - P.n_vertices) + 2];
  double sn, cs;
  sincos(th, &sn, &cs);
  P.trig[2 * static_cast<size_t>(t)] = sn;
  P.trig[2 * static_cast<size_t>(t) + 1] = cs;
}
__global__ void gn_chi2(Params P, int n_partials) {
  __shared__ double scratch[32];
  P = params_at(P, blockIdx.y);
  if (*reinterpret_cast<volatile int*>(P.status) != 0) return;
  double v = 0.0;
  for (int b = threadIdx.x; b < n_partials; b += blockDim.x) v += P.chi2_partial[b];
  const double s = block_sum(v, scratch);
  if (threadIdx.x == 0) P.chi2_out[P.status[1]] = s;  // status[1] = iterations done so far
}
// VertexSE2::oplusImpl (C3); the last kernel of an iteration. One thread per pose scalar: the
// estimates are read and written coalesced, the step x is gathered through vpos (L2-resident).
__global__ void gn_update(Params P) {
  P = params_at(P, blockIdx.y);
  if (*reinterpret_cast<volatile int*>(P.status) != 0) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 3 * P.n_vertices) {
    const int v = t / 3, k = t - 3 * v;
    const int p = P.vpos[v];
    if (p >= 0) {
      const double q = P.poses[t] + P.x[3 * static_cast<size_t>(p) + k];
      P.poses[t] = k == 2 ? normalize_theta(q) : q;
    }
  }
}
__global__ void gn_count_iteration(Params P) {  // one thread per instance
  P = params_at(P, threadIdx.x);
  if (*reinterpret_cast<volatile int*>(P.status) == 0) P.status[1] += 1;
}
__global__ void gn_stamp(unsigned long long* stamps, int k) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  stamps[k] = t;
}

#include "pgo_dd.cuh"

__global__ void __launch_bounds__(kThreads) chi2_only(Params P) {
  __shared__ double scratch[32];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  double chi = 0.0;
  for (int e = tid; e < P.n_edges; e += nthreads) {
    EdgeLin L;
    linearise_edge(P, e, &L);
    chi += edge_chi2(L);
  }
  const double s = block_sum(chi, scratch);
  if (threadIdx.x == 0) P.chi2_partial[blockIdx.x] = s;
}

__global__ void sum_partials(const double* partial, int n, double* out) {
  __shared__ double scratch[32];
  double v = 0.0;
  for (int b = threadIdx.x; b < n; b += blockDim.x) v += partial[b];
  const double s = block_sum(v, scratch);
  if (threadIdx.x == 0) *out = s;
}

__global__ void set_unit_rhs(double* rhs, size_t stride, const int* col_p, int n_cols) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n_cols) return;
  const int k = t / 3, c = t % 3;  // right-hand side 3k + c has a one at scalar row 3 * col_p[k] + c
  rhs[static_cast<size_t>(3 * k + c) * stride + 3 * static_cast<size_t>(col_p[k]) + c] = 1.0;
}

__global__ void gather_blocks(const double* x, size_t stride, const int* row_p, const int* rhs_of,
                              int n_blocks, double* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 9 * n_blocks) return;
  const int k = t / 9, r = (t % 9) / 3, c = t % 3;
  out[t] = x[static_cast<size_t>(3 * rhs_of[k] + c) * stride + 3 * static_cast<size_t>(row_p[k]) + r];
}

// EdgeLabeler::labelEdge for star edges with the gauge fixed.
__global__ void label_star_edges(const double* poses, int gauge, int n, const int* v,
                                 const double* cov, double* meas_out, double* info_out,
                                 int* status) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const SE2d xg = load_pose(poses, gauge), xv = load_pose(poses, v[t]);
  const SE2d xg_inv = se2_inv(xg);
  const SE2d z = se2_mul(xg_inv, xv);  // setMeasurementFromState
  const SE2d zinv = se2_inv(z);
  // sampleUnscented: dim 3, alpha 1e-3, beta 2, lambda = alpha^2 * dim
  const double alpha = 1e-3, beta = 2.0, dim = 3.0;
  const double lam = alpha * alpha * dim;
  const double* c = cov + 9 * static_cast<size_t>(t);
  double a[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) a[i] = c[i] * (dim + lam);
  // lower Cholesky factor of a
  double L[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  bool ok = a[0] > 0.0;
  if (ok) {
    L[0] = sqrt(a[0]);
    L[3] = a[3] / L[0];
    L[6] = a[6] / L[0];
    const double d1 = a[4] - L[3] * L[3];
    ok = d1 > 0.0;
    if (ok) {
      L[4] = sqrt(d1);
      L[7] = (a[7] - L[6] * L[3]) / L[4];
      const double d2 = a[8] - L[6] * L[6] - L[7] * L[7];
      ok = d2 > 0.0;
      if (ok) L[8] = sqrt(d2);
    }
  }
  if (!ok) {
    atomicExch(status, 1);
    return;
  }
  const double wi = 1.0 / (2.0 * (dim + lam));
  const double w0m = lam / (dim + lam);
  const double w0c = w0m + (1.0 - alpha * alpha + beta);
  double errs[7][3], wm[7], wc[7];
  for (int k = 0; k < 7; ++k) {
    double p[3] = {0.0, 0.0, 0.0};
    if (k > 0) {
      const int col = (k - 1) / 2;
      const double sgn = (k - 1) % 2 ? -1.0 : 1.0;
      p[0] = sgn * L[col];
      p[1] = sgn * L[3 + col];
      p[2] = sgn * L[6 + col];
    }
    SE2d pv = xv;  // oplus
    pv.x += p[0];
    pv.y += p[1];
    pv.th = normalize_theta(pv.th + p[2]);
    const SE2d e = se2_mul(zinv, se2_mul(xg_inv, pv));
    errs[k][0] = e.x;
    errs[k][1] = e.y;
    errs[k][2] = e.th;
    wm[k] = k ? wi : w0m;
    wc[k] = k ? wi : w0c;
  }
  double mean[3] = {0.0, 0.0, 0.0};
  for (int k = 0; k < 7; ++k)
    for (int i = 0; i < 3; ++i) mean[i] += wm[k] * errs[k][i];
  double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < 7; ++k)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) s[3 * i + j] += wc[k] * (errs[k][i] - mean[i]) * (errs[k][j] - mean[j]);
  double inv[9];
  if (!spd_inverse(s, inv)) {
    atomicExch(status, 1);
    return;
  }
  meas_out[3 * t] = z.x;
  meas_out[3 * t + 1] = z.y;
  meas_out[3 * t + 2] = z.th;
  for (int i = 0; i < 9; ++i) info_out[9 * static_cast<size_t>(t) + i] = inv[i];
}

template <typename T>
struct Buf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap && p) return cudaSuccess;
    // a buffer that grows again (the keyframe loop adds a vertex per keyframe: every structure array is
    // one element short every time) gets half as much again, up to 64 MB of slack: cudaFree
    // synchronises the device and the pair costs ~0.1 ms per buffer, ~30 buffers per pgo_set_graph
    size_t want = n ? n : 1;
    if (p && cap) want = std::max(want, std::min(cap + cap / 2, cap + (size_t(64) << 20) / sizeof(T)));
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = reserve(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

const int kMaxItersPerCall = 1024;  // capacity of the device-side chi2 record

struct DeviceSolver {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  int grid = 0;  // co-resident CTAs of the cooperative kernels
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t aux = nullptr, aux2 = nullptr, aux3 = nullptr;  // further branches while capturing the iteration graph
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr, ev_join3 = nullptr;
  Params P;
  bool have_structure = false, have_values = false, have_factor = false;
  Buf<int> edge_i, edge_j, vpos, inc_ptr, ff_edges, col_ptr, row_idx, perm_vertex, status, scratch_i;
  Buf<Incidence> inc;
  Buf<int4> winc;
  Buf<int2> winc_ends;
  Buf<int> wg_ptr, wg_info;
  Buf<unsigned long long> stamps;
  double stage_ms[5] = {0, 0, 0, 0, 0};
  Buf<double> poses, meas, info6, M, Dinv, rhs, u, x, chi2_partial, chi2_out, many_rhs, scratch_d, trig;
  // supernodal tables (single-GPU path)
  SNView V;
  // task lists of one stage: everything (single GPU), or this rank's panels [0] and the shared
  // separator panels [1] (domain decomposition)
  struct TaskSet {
    Supernodal::Lists L;   // host copy: level pointers and launch geometry
    Buf<Task> ff, fa, fb, ss, sa, sf, sb;
  } sets[2];
  // [0]: the whole iteration (single GPU) or the local stage; [1]: the shared stage
  cudaGraph_t graph[2] = {nullptr, nullptr};
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  int graph_nodes[2] = {0, 0};
  int lin_blocks = 0;
  int batch = 1;  // problem instances sharing this structure
  Buf<int> colbase, tbl_off, tbl;
  Buf<PanelDesc> pn_desc;
  Buf<SuperDesc> sn_desc;
  Buf<double> diag_scratch;
  Buf<double> Y;
  Buf<double> many_u, many_x;  // substitution vectors of the marginals (u / x never move)
  // domain decomposition
  DDParams D;
  Buf<int> owner;
  Buf<double> pose_x;
  long long exch_first = 0, exch_count = 0;  // doubles, relative to M
};

int dev_create(DeviceSolver** out, int device, void* stream, std::string* err) {
  *out = nullptr;
  int n_dev = 0;
  PGO_CUDA(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) {
    if (err) *err = "no such CUDA device";
    return PGO_ERR_CUDA;
  }
  PGO_CUDA(cudaSetDevice(device));
  int coop = 0;
  PGO_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
  if (!coop) {
    if (err) *err = "device does not support cooperative launches";
    return PGO_ERR_CUDA;
  }
  DeviceSolver* d = new DeviceSolver();
  d->device = device;
  cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, device);
  int per_sm = 0;
  cudaError_t e = cudaSuccess;
  {
    const int cta_bytes = static_cast<int>(sizeof(double) * kCtaSmemDoubles);
    const int warp_bytes = static_cast<int>(sizeof(double) * (std::max(kWarpSmemDoubles, kWarpSubstDoubles) + 1) *
                                            kWarpsPerCta);
    const void* cta_kernels[] = {reinterpret_cast<const void*>(sn_k_factor), reinterpret_cast<const void*>(sn_k_update),
                                 reinterpret_cast<const void*>(sn_k_bwd_tri), reinterpret_cast<const void*>(sn_k_fwd_tri)};
    for (size_t i = 0; i < sizeof(cta_kernels) / sizeof(cta_kernels[0]) && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(cta_kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, cta_bytes);
    const void* warp_kernels[] = {reinterpret_cast<const void*>(sn_k_fused),                                   reinterpret_cast<const void*>(sn_k_bwd_small), reinterpret_cast<const void*>(sn_k_fwd_small)};
    for (size_t i = 0; i < sizeof(warp_kernels) / sizeof(warp_kernels[0]) && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(warp_kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, warp_bytes);
  }
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chi2_only, kThreads, 0);
  if (e == cudaSuccess) {
    per_sm = std::max(1, per_sm);
    d->grid = d->sm_count * per_sm;
    if (stream) {
      d->stream = static_cast<cudaStream_t>(stream);
    } else {
      e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
      d->own_stream = e == cudaSuccess;
    }
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->aux, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->aux2, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->aux3, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_join2, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_join3, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreate(&d->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&d->ev1);
  if (e == cudaSuccess) e = d->status.reserve(4);
  if (e == cudaSuccess) e = d->stamps.reserve(8);
  if (e == cudaSuccess) e = d->chi2_partial.reserve(std::max(d->grid, 4 * d->sm_count));
  if (e != cudaSuccess) {
    if (err) *err = std::string("pgo dev_create: ") + cudaGetErrorString(e);
    dev_destroy(d);
    return PGO_ERR_CUDA;
  }
  *out = d;
  return PGO_OK;
}

void dev_destroy(DeviceSolver* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  if (d->stream) cudaStreamSynchronize(d->stream);
  Buf<int>* ib[] = {&d->edge_i, &d->edge_j, &d->vpos,        &d->inc_ptr, &d->ff_edges,
                    &d->col_ptr, &d->row_idx, &d->perm_vertex, &d->status,  &d->scratch_i};
  for (size_t i = 0; i < sizeof(ib) / sizeof(ib[0]); ++i) ib[i]->release();
  Buf<double>* db[] = {&d->poses, &d->meas, &d->info6, &d->M, &d->Dinv, &d->rhs, &d->u, &d->x,
                       &d->chi2_partial, &d->chi2_out, &d->many_rhs, &d->scratch_d, &d->trig};
  for (size_t i = 0; i < sizeof(db) / sizeof(db[0]); ++i) db[i]->release();
  Buf<int>* sb[] = {&d->colbase, &d->tbl_off, &d->tbl};
  d->pn_desc.release();
  d->sn_desc.release();
  for (size_t i = 0; i < sizeof(sb) / sizeof(sb[0]); ++i) sb[i]->release();
  for (int k = 0; k < 2; ++k) {
    Buf<Task>* tb[] = {&d->sets[k].ff, &d->sets[k].fa, &d->sets[k].fb, &d->sets[k].ss,
                       &d->sets[k].sa, &d->sets[k].sf, &d->sets[k].sb};
    for (size_t i = 0; i < sizeof(tb) / sizeof(tb[0]); ++i) tb[i]->release();
  }
  d->diag_scratch.release();
  d->Y.release();
  d->many_u.release();
  d->many_x.release();
  for (int k = 0; k < 2; ++k) {
    if (d->graph_exec[k]) cudaGraphExecDestroy(d->graph_exec[k]);
    if (d->graph[k]) cudaGraphDestroy(d->graph[k]);
    d->graph_exec[k] = nullptr;
    d->graph[k] = nullptr;
  }
  d->inc.release();
  d->winc.release();
  d->winc_ends.release();
  d->wg_ptr.release();
  d->wg_info.release();
  d->stamps.release();
  d->owner.release();
  d->pose_x.release();
  if (d->ev_fork) cudaEventDestroy(d->ev_fork);
  if (d->ev_join) cudaEventDestroy(d->ev_join);
  if (d->ev_join2) cudaEventDestroy(d->ev_join2);
  if (d->ev_join3) cudaEventDestroy(d->ev_join3);
  if (d->aux) cudaStreamDestroy(d->aux);
  if (d->aux2) cudaStreamDestroy(d->aux2);
  if (d->aux3) cudaStreamDestroy(d->aux3);
  if (d->ev0) cudaEventDestroy(d->ev0);
  if (d->ev1) cudaEventDestroy(d->ev1);
  if (d->own_stream && d->stream) cudaStreamDestroy(d->stream);
  delete d;
}

void* dev_stream(const DeviceSolver* d) { return d->stream; }
uint64_t dev_launches(const DeviceSolver* d) { return d->launches; }
const double* dev_stage_ms(const DeviceSolver* d) { return d->stage_ms; }

int dev_set_structure(DeviceSolver* d, const Symbolic& S, const GraphTables& G, std::string* err) {
  PGO_CUDA(cudaSetDevice(d->device));
  cudaStream_t s = d->stream;
  d->have_structure = d->have_values = d->have_factor = false;
  PGO_CUDA(d->edge_i.upload(G.edge_i, s));
  PGO_CUDA(d->edge_j.upload(G.edge_j, s));
  PGO_CUDA(d->vpos.upload(G.vpos, s));
  PGO_CUDA(d->inc_ptr.upload(G.inc_ptr, s));
  PGO_CUDA(d->inc.upload(G.inc, s));
  PGO_CUDA(d->ff_edges.upload(G.ff_edges, s));
  // warp groups of the linearisation: whole vertices, <= 32 incidences per group (a vertex of higher
  // degree is a group of its own); inside a vertex the entries are ordered by the off-diagonal block
  // they own, so that parallel edges sit next to each other
  std::vector<int4> winc;
  std::vector<int> wg_ptr(1, 0), wg_info;
  {
    winc.reserve(G.inc.size());
    int in_group = 0, flags = 0;
    auto close_group = [&]() {
      if (static_cast<int>(winc.size()) > wg_ptr.back()) {
        wg_ptr.push_back(static_cast<int>(winc.size()));
        wg_info.push_back(flags);
      }
      in_group = 0;
      flags = 0;
    };
    std::vector<int4> mine;
    for (int p = 0; p < S.n; ++p) {
      const int deg = G.inc_ptr[p + 1] - G.inc_ptr[p];
      if (deg == 0) continue;
      mine.clear();
      for (int t = G.inc_ptr[p]; t < G.inc_ptr[p + 1]; ++t) {
        const int edge = G.inc[t].edge_role >> 1, role = G.inc[t].edge_role & 1;
        const int chi = role == 0 || G.vpos[G.edge_i[edge]] < 0 ? 1 : 0;  // every edge's chi2 once
        mine.push_back(make_int4(edge, G.inc[t].pos, p, role | (chi << 1)));
      }
      std::stable_sort(mine.begin(), mine.end(), [](const int4& a, const int4& b) { return a.y < b.y; });
      int dup = 0;
      for (size_t k = 1; k < mine.size(); ++k) dup |= mine[k].y >= 0 && mine[k].y == mine[k - 1].y;
      if (deg > 32 || in_group + deg > 32) close_group();
      winc.insert(winc.end(), mine.begin(), mine.end());
      in_group += deg;
      flags |= dup;
      if (deg > 32) close_group();
    }
    close_group();
    if (wg_info.empty()) {  // no incidences at all: one empty group (it still sums the chi2 of fixed-fixed edges)
      wg_ptr.push_back(0);
      wg_info.push_back(0);
    }
  }
  PGO_CUDA(d->winc.upload(winc, s));
  {
    std::vector<int2> ends(winc.size());
    for (size_t k = 0; k < winc.size(); ++k) ends[k] = make_int2(G.edge_i[winc[k].x], G.edge_j[winc[k].x]);
    PGO_CUDA(d->winc_ends.upload(ends, s));
  }
  PGO_CUDA(d->wg_ptr.upload(wg_ptr, s));
  PGO_CUDA(d->wg_info.upload(wg_info, s));
  PGO_CUDA(d->col_ptr.upload(S.col_ptr, s));
  PGO_CUDA(d->row_idx.upload(S.row_idx, s));
  const Supernodal& N = S.sn;
  PGO_CUDA(d->pn_desc.upload(N.pn, s));
  PGO_CUDA(d->sn_desc.upload(N.sn, s));
  PGO_CUDA(d->colbase.upload(N.colbase, s));
  PGO_CUDA(d->tbl_off.upload(N.tbl_off, s));
  PGO_CUDA(d->tbl.upload(N.tbl, s));
  for (int k = 0; k < 2; ++k) {
    // single GPU: one set with everything; domain decomposition: own panels, then shared panels
    DeviceSolver::TaskSet& ts = d->sets[k];
    ts.L = N.lists(S.world == 1 ? (k == 0 ? Supernodal::kAllOwners : -3) : (k == 0 ? G.rank : -1));
    PGO_CUDA(ts.ff.upload(ts.L.ff, s));
    PGO_CUDA(ts.fa.upload(ts.L.fa, s));
    PGO_CUDA(ts.fb.upload(ts.L.fb, s));
    PGO_CUDA(ts.ss.upload(ts.L.ss, s));
    PGO_CUDA(ts.sa.upload(ts.L.sa, s));
    PGO_CUDA(ts.sf.upload(ts.L.sf, s));
    PGO_CUDA(ts.sb.upload(ts.L.sb, s));
  }
  const size_t scratch_stride = 9 * static_cast<size_t>(N.scratch_blocks) + 9;
  PGO_CUDA(d->diag_scratch.reserve(static_cast<size_t>(d->batch) * scratch_stride));
  std::vector<int> perm_vertex(S.n);
  for (int v = 0; v < G.n_vertices; ++v)
    if (G.vpos[v] >= 0) perm_vertex[G.vpos[v]] = v;
  PGO_CUDA(d->perm_vertex.upload(perm_vertex, s));
  const size_t B = static_cast<size_t>(d->batch);
  PGO_CUDA(d->poses.reserve(B * 3 * static_cast<size_t>(G.n_vertices)));
  PGO_CUDA(d->meas.reserve(B * 3 * static_cast<size_t>(G.n_edges) + 1));
  PGO_CUDA(d->info6.reserve(B * 6 * static_cast<size_t>(G.n_edges) + 1));
  const size_t n_shared = static_cast<size_t>(S.n - S.first_shared);
  // per-instance stride of the factor storage: a multiple of 32 doubles
  const size_t m_stride = (9 * static_cast<size_t>(S.nnzb) + 3 * n_shared + 8 + 31) / 32 * 32;
  PGO_CUDA(d->M.reserve(B * m_stride));
  PGO_CUDA(d->Y.reserve(B * m_stride));
  PGO_CUDA(d->owner.upload(S.owner, s));
  PGO_CUDA(d->pose_x.reserve(3 * static_cast<size_t>(G.n_vertices)));
  PGO_CUDA(d->Dinv.reserve(B * 9 * static_cast<size_t>(S.n)));
  PGO_CUDA(d->rhs.reserve(B * 3 * static_cast<size_t>(S.n)));
  PGO_CUDA(d->u.reserve(B * 3 * static_cast<size_t>(S.n)));
  PGO_CUDA(d->x.reserve(B * 3 * static_cast<size_t>(S.n)));
  PGO_CUDA(cudaStreamSynchronize(s));  // the host vectors may go away
  Params& P = d->P;
  P.n_vertices = G.n_vertices;
  P.n_edges = G.n_edges;
  P.n = S.n;
  P.edge_i = d->edge_i.p;
  P.edge_j = d->edge_j.p;
  P.vpos = d->vpos.p;
  P.inc_ptr = d->inc_ptr.p;
  P.inc = d->inc.p;
  P.ff_edges = d->ff_edges.p;
  P.n_ff = static_cast<int>(G.ff_edges.size());
  P.winc = d->winc.p;
  P.winc_ends = d->winc_ends.p;
  P.wg_ptr = d->wg_ptr.p;
  P.wg_info = d->wg_info.p;
  P.n_wgroups = static_cast<int>(wg_info.size());
  P.poses = d->poses.p;
  P.meas = d->meas.p;
  P.info6 = d->info6.p;
  P.nnzb = S.nnzb;
  P.col_ptr = d->col_ptr.p;
  P.row_idx = d->row_idx.p;
  P.perm_vertex = d->perm_vertex.p;
  P.M = d->M.p;
  P.Dinv = d->Dinv.p;
  P.rhs = d->rhs.p;
  P.u = d->u.p;
  P.x = d->x.p;
  P.chi2_partial = d->chi2_partial.p;
  P.chi2_out = nullptr;
  P.status = d->status.p;
  P.stamps = d->stamps.p;
  SNView& V = d->V;
  V.row_idx = d->row_idx.p;
  V.pn = d->pn_desc.p;
  V.sn = d->sn_desc.p;
  V.colbase = d->colbase.p;
  V.tbl_off = d->tbl_off.p;
  V.tbl = d->tbl.p;
  V.M = d->M.p;
  V.Y = d->Y.p;
  V.Dinv = d->Dinv.p;
  V.z = d->rhs.p;
  V.u = d->u.p;
  V.x = d->x.p;
  V.scratch = d->diag_scratch.p;
  V.status = d->status.p;
  V.s_M = static_cast<long long>(m_stride);
  V.s_Dinv = 9LL * S.n;
  V.s_vec = 3LL * S.n;
  V.s_scratch = static_cast<long long>(scratch_stride);
  V.s_status = 4;
  d->lin_blocks = (P.n_wgroups + kLinThreads / 32 - 1) / (kLinThreads / 32);  // one warp per group
  PGO_CUDA(d->chi2_out.reserve(B * kMaxItersPerCall));
  PGO_CUDA(d->chi2_partial.reserve(std::max(B * d->lin_blocks * (kLinThreads / 32), static_cast<size_t>(d->grid))));
  PGO_CUDA(d->status.reserve(4 * B));
  P.chi2_partial = d->chi2_partial.p;
  P.status = d->status.p;
  V.status = d->status.p;
  P.s_poses = 3LL * G.n_vertices;
  P.s_meas = 3LL * G.n_edges;
  P.s_info = 6LL * G.n_edges;
  P.s_M = V.s_M;
  P.s_Dinv = V.s_Dinv;
  P.s_vec = V.s_vec;
  P.s_partial = d->lin_blocks * (kLinThreads / 32);
  P.s_chi2 = kMaxItersPerCall;
  P.s_trig = 2LL * (static_cast<long long>(G.n_vertices) + G.n_edges);
  PGO_CUDA(d->trig.reserve(B * static_cast<size_t>(P.s_trig)));
  P.trig = d->trig.p;
  for (int k = 0; k < 2; ++k) {
    if (d->graph_exec[k]) cudaGraphExecDestroy(d->graph_exec[k]);
    if (d->graph[k]) cudaGraphDestroy(d->graph[k]);
    d->graph_exec[k] = nullptr;
    d->graph[k] = nullptr;
  }
  DDParams& D = d->D;
  D.owner = d->owner.p;
  D.rank = G.rank;
  D.world = S.world;
  D.first_shared = S.first_shared;
  D.tail = d->M.p + 9 * static_cast<size_t>(S.nnzb);
  D.pose_x = d->pose_x.p;
  d->exch_first = 9LL * (S.n > S.first_shared ? S.col_ptr[S.first_shared] : S.nnzb);
  d->exch_count = 9LL * S.nnzb + 3LL * static_cast<long long>(n_shared) + 2 - d->exch_first;
  d->have_structure = true;
  return PGO_OK;
}

int dev_set_batch(DeviceSolver* d, int batch, std::string* err) {
  if (batch < 1 || batch > 1024) {
    if (err) *err = "batch must be in [1, 1024]";
    return PGO_ERR_ARG;
  }
  if (batch != d->batch) d->have_structure = d->have_values = d->have_factor = false;
  d->batch = batch;
  return PGO_OK;
}
int dev_batch(const DeviceSolver* d) { return d->batch; }

// inst = -1: every instance of the batch gets the same values
int dev_upload(DeviceSolver* d, int inst, const double* poses, const double* meas,
               const double* info6, std::string* err) {
  if (!d->have_structure) {
    if (err) *err = "pgo_set_graph has not been called";
    return PGO_ERR_ARG;
  }
  if (inst < -1 || inst >= d->batch) {
    if (err) *err = "no such instance";
    return PGO_ERR_ARG;
  }
  PGO_CUDA(cudaSetDevice(d->device));
  const Params& P = d->P;
  for (int b = (inst < 0 ? 0 : inst); b < (inst < 0 ? d->batch : inst + 1); ++b) {
    PGO_CUDA(cudaMemcpyAsync(d->poses.p + b * P.s_poses, poses, 3 * sizeof(double) * P.n_vertices,
                             cudaMemcpyHostToDevice, d->stream));
    if (P.n_edges) {
      PGO_CUDA(cudaMemcpyAsync(d->meas.p + b * P.s_meas, meas, 3 * sizeof(double) * P.n_edges,
                               cudaMemcpyHostToDevice, d->stream));
      PGO_CUDA(cudaMemcpyAsync(d->info6.p + b * P.s_info, info6, 6 * sizeof(double) * P.n_edges,
                               cudaMemcpyHostToDevice, d->stream));
    }
  }
  PGO_CUDA(cudaStreamSynchronize(d->stream));
  if (inst <= 0) d->have_values = true;  // instance 0 (or all) given: the solver is usable
  d->have_factor = false;
  return PGO_OK;
}

int dev_set_poses(DeviceSolver* d, int inst, const double* poses, std::string* err) {
  if (!d->have_values) {
    if (err) *err = "pgo_upload has not been called";
    return PGO_ERR_ARG;
  }
  if (inst < -1 || inst >= d->batch) {
    if (err) *err = "no such instance";
    return PGO_ERR_ARG;
  }
  PGO_CUDA(cudaSetDevice(d->device));
  for (int b = (inst < 0 ? 0 : inst); b < (inst < 0 ? d->batch : inst + 1); ++b)
    PGO_CUDA(cudaMemcpyAsync(d->poses.p + b * d->P.s_poses, poses, 3 * sizeof(double) * d->P.n_vertices,
                             cudaMemcpyHostToDevice, d->stream));
  PGO_CUDA(cudaStreamSynchronize(d->stream));
  return PGO_OK;
}

int dev_get_poses(DeviceSolver* d, int inst, double* poses, std::string* err) {
  if (!d->have_values) {
    if (err) *err = "pgo_upload has not been called";
    return PGO_ERR_ARG;
  }
  if (inst < 0 || inst >= d->batch) {
    if (err) *err = "no such instance";
    return PGO_ERR_ARG;
  }
  PGO_CUDA(cudaSetDevice(d->device));
  PGO_CUDA(cudaMemcpyAsync(poses, d->poses.p + inst * d->P.s_poses, 3 * sizeof(double) * d->P.n_vertices,
                           cudaMemcpyDeviceToHost, d->stream));
  PGO_CUDA(cudaStreamSynchronize(d->stream));
  return PGO_OK;
}

// ---- recording the phases of an iteration on the solver's stream (captured into CUDA graphs) -----
// Factorisation + forward substitution of one task set, bottom-up over the panel levels.
static int enqueue_factor(DeviceSolver* d, const DeviceSolver::TaskSet& ts, int* nodes, std::string* err) {
  cudaStream_t st = d->stream;
  const SNView V = d->V;
  const Supernodal::Lists& L = ts.L;
  const int B = d->batch;
  for (int l = 0; l < L.n_plevels; ++l) {
    const int n_fa = L.fa_ptr[l + 1] - L.fa_ptr[l], n_ff = L.ff_ptr[l + 1] - L.ff_ptr[l],
              n_fb = L.fb_ptr[l + 1] - L.fb_ptr[l];
    // the level's big panels (CTA tasks), its small panels (warp tasks) and the few small panels
    // with many rows (warp tasks with a large shared-memory stride) are independent: up to three
    // branches of the graph
    const int n_ffl = L.ff_large[l], n_ffs = n_ff - n_ffl;
    cudaStream_t s_small = st, s_large = st;
    if (n_ffs && (n_fa || n_ffl)) s_small = d->aux;
    if (n_ffl && n_fa) s_large = d->aux2;
    if (s_small != st || s_large != st) PGO_CUDA(cudaEventRecord(d->ev_fork, st));
    if (s_small != st) PGO_CUDA(cudaStreamWaitEvent(s_small, d->ev_fork, 0));
    if (s_large != st) PGO_CUDA(cudaStreamWaitEvent(s_large, d->ev_fork, 0));
    if (n_ffl) {
      const int stride = (L.ff_smem[l] + 1) & ~1;
      sn_k_fused<<<dim3((n_ffl + kWarpsPerCta - 1) / kWarpsPerCta, B), 32 * kWarpsPerCta,
                   sizeof(double) * stride * kWarpsPerCta, s_large>>>(V, ts.ff.p + L.ff_ptr[l], n_ffl, stride);
      ++*nodes;
    }
    if (n_ffs) {
      const int stride = (L.ff_smem_small[l] + 1) & ~1;
      sn_k_fused<<<dim3((n_ffs + kWarpsPerCta - 1) / kWarpsPerCta, B), 32 * kWarpsPerCta,
                   sizeof(double) * stride * kWarpsPerCta, s_small>>>(V, ts.ff.p + L.ff_ptr[l] + n_ffl, n_ffs,
                                                                      stride);
      ++*nodes;
    }
    // wide panels on the main branch, narrow ones (less shared memory: more CTAs per SM) beside them
    const int n_fal = L.fa_large[l], n_fas = n_fa - n_fal;
    cudaStream_t s_fas = st;
    if (n_fas && n_fal) {
      if (s_small == st && s_large == st) PGO_CUDA(cudaEventRecord(d->ev_fork, st));
      PGO_CUDA(cudaStreamWaitEvent(d->aux3, d->ev_fork, 0));
      s_fas = d->aux3;
    }
    if (n_fal) {
      sn_k_factor<<<dim3(n_fal, B), kCtaThreads, sizeof(double) * L.fa_smem[l], st>>>(V, ts.fa.p + L.fa_ptr[l]);
      ++*nodes;
    }
    if (n_fas) {
      sn_k_factor<<<dim3(n_fas, B), kCtaThreads / 2, sizeof(double) * L.fa_smem_small[l], s_fas>>>(
          V, ts.fa.p + L.fa_ptr[l] + n_fal);
      ++*nodes;
    }
    if (s_fas != st) {
      PGO_CUDA(cudaEventRecord(d->ev_join3, s_fas));
      PGO_CUDA(cudaStreamWaitEvent(st, d->ev_join3, 0));
    }
    if (s_small != st) {
      PGO_CUDA(cudaEventRecord(d->ev_join, s_small));
      PGO_CUDA(cudaStreamWaitEvent(st, d->ev_join, 0));
    }
    if (s_large != st) {
      PGO_CUDA(cudaEventRecord(d->ev_join2, s_large));
      PGO_CUDA(cudaStreamWaitEvent(st, d->ev_join2, 0));
    }
    if (n_fb) {
      sn_k_update<<<dim3(n_fb, B), kUpdateThreads, sizeof(double) * L.fb_smem[l], st>>>(V, ts.fb.p + L.fb_ptr[l]);
      ++*nodes;
    }
  }
  PGO_CUDA(cudaGetLastError());
  return PGO_OK;
}

// Backward substitution of one task set, top-down over the supernode levels.
static int enqueue_backward(DeviceSolver* d, const DeviceSolver::TaskSet& ts, const SNView& V, int B,
                            int* nodes, std::string* err) {
  cudaStream_t st = d->stream;
  const Supernodal::Lists& L = ts.L;
  for (int l = L.n_slevels - 1; l >= 0; --l) {
    const int n_sa = L.sa_ptr[l + 1] - L.sa_ptr[l], n_ss = L.ss_ptr[l + 1] - L.ss_ptr[l],
              n_sb = L.sb_ptr[l + 1] - L.sb_ptr[l];
    // narrow supernodes (whole, one warp each) next to the wide ones (row gather, then the
    // triangular part): two branches
    cudaStream_t side = st;
    if (n_ss && (n_sb || n_sa)) {
      PGO_CUDA(cudaEventRecord(d->ev_fork, st));
      PGO_CUDA(cudaStreamWaitEvent(d->aux, d->ev_fork, 0));
      side = d->aux;
    }
    const int n_ssl = L.ss_large[l];
    if (n_ssl) {  // single-panel supernodes of 5 .. 16 columns
      const int stride = (L.ss_smem[l] + 1) & ~1;
      sn_k_bwd_small<<<dim3((n_ssl + kWarpsPerCta - 1) / kWarpsPerCta, B), 32 * kWarpsPerCta,
                       sizeof(double) * stride * kWarpsPerCta, side>>>(V, ts.ss.p + L.ss_ptr[l], n_ssl, stride);
      ++*nodes;
    }
    if (n_ss > n_ssl) {
      const int stride = (L.ss_smem_small[l] + 1) & ~1;
      sn_k_bwd_small<<<dim3((n_ss - n_ssl + kWarpsPerCta - 1) / kWarpsPerCta, B), 32 * kWarpsPerCta,
                       sizeof(double) * stride * kWarpsPerCta, side>>>(V, ts.ss.p + L.ss_ptr[l] + n_ssl,
                                                                       n_ss - n_ssl, stride);
      ++*nodes;
    }
    if (n_sb) {
      sn_k_bwd_rows<<<dim3((n_sb + kWarpsPerCta - 1) / kWarpsPerCta, B), 32 * kWarpsPerCta, 0, st>>>(
          V, ts.sb.p + L.sb_ptr[l], n_sb);
      ++*nodes;
    }
    if (n_sa) {
      sn_k_bwd_tri<<<dim3(n_sa, B), kCtaThreads, sizeof(double) * L.sa_smem[l], st>>>(V, ts.sa.p + L.sa_ptr[l]);
      ++*nodes;
    }
    if (side != st) {
      PGO_CUDA(cudaEventRecord(d->ev_join, side));
      PGO_CUDA(cudaStreamWaitEvent(st, d->ev_join, 0));
    }
  }
  PGO_CUDA(cudaGetLastError());
  return PGO_OK;
}

// Stand-alone forward substitution of B right-hand sides (marginals), bottom-up.
static int enqueue_forward(DeviceSolver* d, const DeviceSolver::TaskSet& ts, const SNView& V, int B,
                           int* nodes, std::string* err) {
  cudaStream_t st = d->stream;
  const Supernodal::Lists& L = ts.L;
  for (int l = 0; l < L.n_slevels; ++l) {
    const int n_sa = L.sa_ptr[l + 1] - L.sa_ptr[l], n_ss = L.ss_ptr[l + 1] - L.ss_ptr[l],
              n_sf = L.sf_ptr[l + 1] - L.sf_ptr[l];
    const int n_ssl = L.ss_large[l];
    if (n_ssl) {
      const int stride = (L.ss_smem[l] + 1) & ~1;
      sn_k_fwd_small<<<dim3((n_ssl + kWarpsPerCta - 1) / kWarpsPerCta, B), 32 * kWarpsPerCta,
                       sizeof(double) * stride * kWarpsPerCta, st>>>(V, ts.ss.p + L.ss_ptr[l], n_ssl, stride);
      ++*nodes;
    }
    if (n_ss > n_ssl) {
      const int stride = (L.ss_smem_small[l] + 1) & ~1;
      sn_k_fwd_small<<<dim3((n_ss - n_ssl + kWarpsPerCta - 1) / kWarpsPerCta, B), 32 * kWarpsPerCta,
                       sizeof(double) * stride * kWarpsPerCta, st>>>(V, ts.ss.p + L.ss_ptr[l] + n_ssl, n_ss - n_ssl,
                                                                     stride);
      ++*nodes;
    }
    if (n_sa) {
      sn_k_fwd_tri<<<dim3(n_sa, B), kCtaThreads, sizeof(double) * L.sa_smem[l], st>>>(V, ts.sa.p + L.sa_ptr[l]);
      ++*nodes;
    }
    if (n_sf) {
      sn_k_fwd_rows<<<dim3((n_sf + kWarpsPerCta - 1) / kWarpsPerCta, B), 32 * kWarpsPerCta, 0, st>>>(
          V, ts.sf.p + L.sf_ptr[l], n_sf);
      ++*nodes;
    }
  }
  PGO_CUDA(cudaGetLastError());
  return PGO_OK;
}

// stage 0: the whole iteration (single GPU) or the local stage of the domain decomposition;
// stage 1: the shared stage (after the all-reduce)
static int enqueue_stage(DeviceSolver* d, int stage, std::string* err) {
  cudaStream_t st = d->stream;
  Params P = d->P;
  P.chi2_out = d->chi2_out.p;
  const int B = d->batch;
  const bool dd = d->D.world > 1;
  int nodes = 0, rc = PGO_OK;
  if (stage == 0) {
    gn_stamp<<<1, 1, 0, st>>>(d->stamps.p, 0);
    // zero the factor storage (fill positions must start at 0; with a domain decomposition also
    // the exchange tail behind it) and the backward accumulators
    PGO_CUDA(cudaMemsetAsync(P.M, 0, sizeof(double) * static_cast<size_t>(P.s_M) * B, st));
    PGO_CUDA(cudaMemsetAsync(P.x, 0, sizeof(double) * static_cast<size_t>(P.s_vec) * B, st));
    if (!dd) PGO_CUDA(cudaMemsetAsync(P.rhs, 0, sizeof(double) * static_cast<size_t>(P.s_vec) * B, st));
    if (dd) gn_dd_linearise<<<d->lin_blocks, kThreads, 0, st>>>(P, d->D);
    else {
      gn_trig<<<dim3((P.n_vertices + P.n_edges + 255) / 256, B), 256, 0, st>>>(P);
      gn_linearise<<<dim3(d->lin_blocks, B), kLinThreads, 0, st>>>(P);
      ++nodes;
    }
    if (!dd) gn_chi2<<<dim3(1, B), 256, 0, st>>>(P, d->lin_blocks * (kLinThreads / 32));
    gn_stamp<<<1, 1, 0, st>>>(d->stamps.p, 1);
    nodes += 6;
    rc = enqueue_factor(d, d->sets[0], &nodes, err);
    if (rc != PGO_OK) return rc;
    if (dd) {
      gn_dd_pack<<<1, 256, 0, st>>>(P, d->D, d->lin_blocks);
      ++nodes;
    }
    gn_stamp<<<1, 1, 0, st>>>(d->stamps.p, 2);
  }
  if (stage == 1 || !dd) {
    if (dd) {
      gn_dd_unpack<<<1, 256, 0, st>>>(P, d->D);
      ++nodes;
      rc = enqueue_factor(d, d->sets[1], &nodes, err);
      if (rc != PGO_OK) return rc;
    }
    gn_stamp<<<1, 1, 0, st>>>(d->stamps.p, 3);
    if (dd) {
      rc = enqueue_backward(d, d->sets[1], d->V, d->batch, &nodes, err);
      if (rc != PGO_OK) return rc;
    }
    rc = enqueue_backward(d, d->sets[0], d->V, d->batch, &nodes, err);
    if (rc != PGO_OK) return rc;
    gn_stamp<<<1, 1, 0, st>>>(d->stamps.p, 4);
    if (dd) gn_dd_update<<<std::max(1, (P.n + 255) / 256), 256, 0, st>>>(P, d->D);
    else gn_update<<<dim3(std::max(1, (3 * P.n_vertices + 255) / 256), B), 256, 0, st>>>(P);
    gn_count_iteration<<<1, B, 0, st>>>(P);
    gn_stamp<<<1, 1, 0, st>>>(d->stamps.p, 5);
    nodes += 6;
  }
  d->graph_nodes[stage] = nodes;
  PGO_CUDA(cudaGetLastError());
  return PGO_OK;
}

static int build_stage_graph(DeviceSolver* d, int stage, std::string* err) {
  if (d->graph_exec[stage]) cudaGraphExecDestroy(d->graph_exec[stage]);
  if (d->graph[stage]) cudaGraphDestroy(d->graph[stage]);
  d->graph_exec[stage] = nullptr;
  d->graph[stage] = nullptr;
  PGO_CUDA(cudaStreamBeginCapture(d->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = enqueue_stage(d, stage, err);
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(d->stream, &g);
  if (rc != PGO_OK) {
    if (g) cudaGraphDestroy(g);
    return rc;
  }
  PGO_CUDA(e);
  d->graph[stage] = g;
  PGO_CUDA(cudaGraphInstantiate(&d->graph_exec[stage], d->graph[stage], 0));
  return PGO_OK;
}

// chi2_out: [batch][n_iters] (may be null); iters_done: [batch].
int dev_iterate(DeviceSolver* d, int n_iters, double* chi2_out, int* iters_done, float* ms,
                std::string* err) {
  if (!d->have_values) {
    if (err) *err = "pgo_upload has not been called";
    return PGO_ERR_ARG;
  }
  PGO_CUDA(cudaSetDevice(d->device));
  const int B = d->batch;
  for (int b = 0; b < B; ++b) iters_done[b] = 0;
  *ms = 0.f;
  if (n_iters <= 0) return PGO_OK;
  if (d->D.world != 1) {
    if (err) *err = "this solver was analysed for a domain decomposition: use the pgo_dd_* calls";
    return PGO_ERR_ARG;
  }
  if (!d->graph_exec[0]) {
    const int rc = build_stage_graph(d, 0, err);
    if (rc != PGO_OK) return rc;
  }
  std::vector<int> status(4 * static_cast<size_t>(B), 0);
  std::vector<double> chi2(chi2_out ? static_cast<size_t>(B) * kMaxItersPerCall : 0);
  bool failed = false;
  PGO_CUDA(cudaEventRecord(d->ev0, d->stream));
  for (int first = 0; first < n_iters && !failed; first += kMaxItersPerCall) {
    const int chunk = std::min(kMaxItersPerCall, n_iters - first);
    PGO_CUDA(cudaMemsetAsync(d->status.p, 0, 4 * sizeof(int) * B, d->stream));
    PGO_CUDA(cudaMemsetAsync(d->chi2_out.p, 0, sizeof(double) * B * kMaxItersPerCall, d->stream));
    for (int it = 0; it < chunk; ++it) PGO_CUDA(cudaGraphLaunch(d->graph_exec[0], d->stream));
    d->launches += static_cast<uint64_t>(chunk) * d->graph_nodes[0];
    if (first + chunk >= n_iters) PGO_CUDA(cudaEventRecord(d->ev1, d->stream));
    PGO_CUDA(cudaMemcpyAsync(status.data(), d->status.p, 4 * sizeof(int) * B, cudaMemcpyDeviceToHost,
                             d->stream));
    if (chi2_out)
      PGO_CUDA(cudaMemcpyAsync(chi2.data(), d->chi2_out.p, sizeof(double) * B * kMaxItersPerCall,
                               cudaMemcpyDeviceToHost, d->stream));
    PGO_CUDA(cudaStreamSynchronize(d->stream));
    for (int b = 0; b < B; ++b) {
      iters_done[b] += status[4 * b + 1];
      failed = failed || status[4 * b] != 0;
      if (chi2_out)
        for (int it = 0; it < chunk; ++it)
          chi2_out[static_cast<size_t>(b) * n_iters + first + it] = chi2[static_cast<size_t>(b) * kMaxItersPerCall + it];
    }
  }
  if (failed) PGO_CUDA(cudaEventRecord(d->ev1, d->stream));
  unsigned long long st[6] = {0, 0, 0, 0, 0, 0};
  PGO_CUDA(cudaMemcpyAsync(st, d->stamps.p, sizeof st, cudaMemcpyDeviceToHost, d->stream));
  PGO_CUDA(cudaStreamSynchronize(d->stream));
  for (int k = 0; k < 5; ++k) d->stage_ms[k] = st[k + 1] > st[k] ? (st[k + 1] - st[k]) * 1e-6 : 0.0;
  PGO_CUDA(cudaEventElapsedTime(ms, d->ev0, d->ev1));
  // a failed iteration has zeroed and partly overwritten the factor of the last good one: no
  // marginals / edge labels from it (they would come back as zeros)
  d->have_factor = !failed && iters_done[0] > 0;
  if (failed) {
    if (err) *err = "H is not positive definite (a 3x3 pivot failed): is a vertex fixed?";
    return PGO_ERR_NUMERIC;
  }
  return PGO_OK;
}

int dev_dd_begin(DeviceSolver* d, int n_iters, std::string* err) {
  if (!d->have_values) {
    if (err) *err = "pgo_upload has not been called";
    return PGO_ERR_ARG;
  }
  if (n_iters > kMaxItersPerCall) {  // the stage graphs hold chi2_out by address: its capacity is fixed
    if (err) *err = "at most 1024 iterations per domain-decomposed call";
    return PGO_ERR_ARG;
  }
  PGO_CUDA(cudaSetDevice(d->device));
  PGO_CUDA(cudaMemsetAsync(d->status.p, 0, 4 * sizeof(int), d->stream));
  PGO_CUDA(cudaMemsetAsync(d->chi2_out.p, 0, std::max(n_iters, 1) * sizeof(double), d->stream));
  PGO_CUDA(cudaEventRecord(d->ev0, d->stream));
  return PGO_OK;
}

static int dd_launch_stage(DeviceSolver* d, int stage, std::string* err) {
  PGO_CUDA(cudaSetDevice(d->device));
  if (d->D.world < 2) {
    if (err) *err = "not a domain-decomposed solver (pgo_set_partition first)";
    return PGO_ERR_ARG;
  }
  if (!d->graph_exec[stage]) {
    const int rc = build_stage_graph(d, stage, err);
    if (rc != PGO_OK) return rc;
  }
  PGO_CUDA(cudaGraphLaunch(d->graph_exec[stage], d->stream));
  d->launches += d->graph_nodes[stage];
  return PGO_OK;
}

int dev_dd_local(DeviceSolver* d, std::string* err) { return dd_launch_stage(d, 0, err); }

void dev_dd_exchange(DeviceSolver* d, void** ptr, long long* n_doubles) {
  *ptr = d->M.p + d->exch_first;
  *n_doubles = d->exch_count;
}

int dev_dd_shared(DeviceSolver* d, std::string* err) { return dd_launch_stage(d, 1, err); }

int dev_dd_end(DeviceSolver* d, int n_iters, double* chi2_out, int* iters_done, float* ms,
               std::string* err) {
  PGO_CUDA(cudaSetDevice(d->device));
  PGO_CUDA(cudaEventRecord(d->ev1, d->stream));
  int status[4] = {0, 0, 0, 0};
  PGO_CUDA(cudaMemcpyAsync(status, d->status.p, sizeof status, cudaMemcpyDeviceToHost, d->stream));
  if (chi2_out && n_iters > 0)
    PGO_CUDA(cudaMemcpyAsync(chi2_out, d->chi2_out.p, n_iters * sizeof(double),
                             cudaMemcpyDeviceToHost, d->stream));
  PGO_CUDA(cudaStreamSynchronize(d->stream));
  PGO_CUDA(cudaEventElapsedTime(ms, d->ev0, d->ev1));
  *iters_done = status[1];
  d->have_factor = status[1] > 0 && !status[0];
  if (status[0]) {
    if (err) *err = "H is not positive definite (a 3x3 pivot failed): is a vertex fixed?";
    return PGO_ERR_NUMERIC;
  }
  return PGO_OK;
}

int dev_dd_pose_exchange(DeviceSolver* d, void** ptr, long long* n_doubles, std::string* err) {
  PGO_CUDA(cudaSetDevice(d->device));
  dd_mask_poses<<<(d->P.n_vertices + 255) / 256, 256, 0, d->stream>>>(d->P, d->D);
  d->launches++;
  PGO_CUDA(cudaGetLastError());
  *ptr = d->pose_x.p;
  *n_doubles = 3LL * d->P.n_vertices;
  return PGO_OK;
}

int dev_dd_pose_commit(DeviceSolver* d, std::string* err) {
  PGO_CUDA(cudaSetDevice(d->device));
  PGO_CUDA(cudaMemcpyAsync(d->poses.p, d->pose_x.p, 3 * sizeof(double) * d->P.n_vertices,
                           cudaMemcpyDeviceToDevice, d->stream));
  return PGO_OK;
}

int dev_chi2(DeviceSolver* d, double* chi2, std::string* err) {
  if (!d->have_values) {
    if (err) *err = "pgo_upload has not been called";
    return PGO_ERR_ARG;
  }
  PGO_CUDA(cudaSetDevice(d->device));
  PGO_CUDA(d->scratch_d.reserve(1));
  const int blocks = std::min(d->grid, std::max(1, (d->P.n_edges + kThreads - 1) / kThreads));
  chi2_only<<<blocks, kThreads, 0, d->stream>>>(d->P);
  sum_partials<<<1, 256, 0, d->stream>>>(d->chi2_partial.p, blocks, d->scratch_d.p);
  d->launches += 2;
  PGO_CUDA(cudaGetLastError());
  PGO_CUDA(cudaMemcpyAsync(chi2, d->scratch_d.p, sizeof(double), cudaMemcpyDeviceToHost, d->stream));
  PGO_CUDA(cudaStreamSynchronize(d->stream));
  return PGO_OK;
}

int dev_marginals(DeviceSolver* d, int n, const int* col_p, const int* row_p, double* cov_out,
                  std::string* err) {
  if (!d->have_factor) {
    if (err) *err = "no factorisation: call pgo_iterate first (computeMarginals uses its H)";
    return PGO_ERR_ARG;
  }
  if (n <= 0) return PGO_OK;
  if (d->D.world != 1) {
    if (err) *err = "marginals need the whole factor: not available on a domain-decomposed solver";
    return PGO_ERR_ARG;
  }
  PGO_CUDA(cudaSetDevice(d->device));
  // distinct columns -> right-hand sides
  std::vector<int> cols(col_p, col_p + n);
  std::sort(cols.begin(), cols.end());
  cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
  std::vector<int> rhs_of(n);
  for (int k = 0; k < n; ++k)
    rhs_of[k] = static_cast<int>(std::lower_bound(cols.begin(), cols.end(), col_p[k]) - cols.begin());
  const size_t stride = 3 * static_cast<size_t>(d->P.n);
  // Columns per pass (3 right-hand sides each, solved as batch instances of the substitution kernels).
  // On the graphs of the keyframe loop (a few hundred poses) a pass is bound by its ~50 dependent
  // launches, not by its work, and the covariance gate asks for a block per candidate vertex
  // (graph_manipulator.cpp:134-142): as many columns per pass as three vector sets of 256 MB allow,
  // at most 64 (measured on the 884-keyframe replay, ms per request: 9.1 with 16 columns per pass, 4.5
  // with 64, 5.8 with 128, 10.5 with 256).
  const size_t cap_cols = std::max<size_t>(16, (256u << 20) / (3 * stride * sizeof(double) + 1));
  const int batch_cols = static_cast<int>(std::min<size_t>(std::min<size_t>(64, cap_cols), std::max<size_t>(cols.size(), 1)));
  PGO_CUDA(d->many_rhs.reserve(3 * batch_cols * stride));
  // the marginals' substitution vectors are separate: the iteration graph holds u / x by address
  PGO_CUDA(d->many_u.reserve(3 * batch_cols * stride));
  PGO_CUDA(d->many_x.reserve(3 * batch_cols * stride));
  PGO_CUDA(d->scratch_i.reserve(3 * static_cast<size_t>(n) + batch_cols));
  PGO_CUDA(d->scratch_d.reserve(9 * static_cast<size_t>(n) + 1));
  std::vector<double> out_all(9 * static_cast<size_t>(n));
  for (size_t c0 = 0; c0 < cols.size(); c0 += batch_cols) {
    const int nc = static_cast<int>(std::min<size_t>(batch_cols, cols.size() - c0));
    // requests served by this batch
    std::vector<int> idx, rows, local_rhs;
    for (int k = 0; k < n; ++k)
      if (rhs_of[k] >= static_cast<int>(c0) && rhs_of[k] < static_cast<int>(c0) + nc) {
        idx.push_back(k);
        rows.push_back(row_p[k]);
        local_rhs.push_back(rhs_of[k] - static_cast<int>(c0));
      }
    const int nb = static_cast<int>(idx.size());
    int* d_cols = d->scratch_i.p;
    int* d_rows = d_cols + batch_cols;
    int* d_rhs_of = d_rows + n;
    PGO_CUDA(cudaMemcpyAsync(d_cols, cols.data() + c0, nc * sizeof(int), cudaMemcpyHostToDevice, d->stream));
    PGO_CUDA(cudaMemcpyAsync(d_rows, rows.data(), nb * sizeof(int), cudaMemcpyHostToDevice, d->stream));
    PGO_CUDA(cudaMemcpyAsync(d_rhs_of, local_rhs.data(), nb * sizeof(int), cudaMemcpyHostToDevice, d->stream));
    PGO_CUDA(cudaMemsetAsync(d->many_rhs.p, 0, 3 * nc * stride * sizeof(double), d->stream));
    set_unit_rhs<<<(3 * nc + 127) / 128, 128, 0, d->stream>>>(d->many_rhs.p, stride, d_cols, nc);
    // the right-hand sides play the role of batch instances that share one factor
    SNView V = d->V;
    V.z = d->many_rhs.p;
    V.u = d->many_u.p;
    V.x = d->many_x.p;
    V.s_M = V.s_Dinv = V.s_scratch = 0;
    V.s_vec = static_cast<long long>(stride);
    V.s_status = 0;
    const int nrhs = 3 * nc;
    PGO_CUDA(cudaMemsetAsync(d->many_x.p, 0, nrhs * stride * sizeof(double), d->stream));
    int nodes = 0;
    int rc = enqueue_forward(d, d->sets[0], V, nrhs, &nodes, err);
    if (rc == PGO_OK) rc = enqueue_backward(d, d->sets[0], V, nrhs, &nodes, err);
    if (rc != PGO_OK) return rc;
    d->launches += nodes;
    gather_blocks<<<(9 * nb + 127) / 128, 128, 0, d->stream>>>(d->many_x.p, stride, d_rows, d_rhs_of, nb,
                                                              d->scratch_d.p);
    d->launches += 2;
    PGO_CUDA(cudaGetLastError());
    std::vector<double> part(9 * static_cast<size_t>(nb));
    PGO_CUDA(cudaMemcpyAsync(part.data(), d->scratch_d.p, part.size() * sizeof(double),
                             cudaMemcpyDeviceToHost, d->stream));
    PGO_CUDA(cudaStreamSynchronize(d->stream));
    for (int t = 0; t < nb; ++t)
      std::copy(part.begin() + 9 * t, part.begin() + 9 * t + 9, out_all.begin() + 9 * idx[t]);
  }
  std::copy(out_all.begin(), out_all.end(), cov_out);
  return PGO_OK;
}

int dev_label_star(DeviceSolver* d, int gauge_vertex, int n, const int* v, const double* cov,
                   double* meas_out, double* info_out, std::string* err) {
  if (n <= 0) return PGO_OK;
  PGO_CUDA(cudaSetDevice(d->device));
  PGO_CUDA(d->scratch_i.reserve(static_cast<size_t>(n) + 4));
  PGO_CUDA(d->scratch_d.reserve(21 * static_cast<size_t>(n)));
  double* d_cov = d->scratch_d.p;
  double* d_meas = d_cov + 9 * static_cast<size_t>(n);
  double* d_info = d_meas + 3 * static_cast<size_t>(n);
  PGO_CUDA(cudaMemcpyAsync(d->scratch_i.p, v, n * sizeof(int), cudaMemcpyHostToDevice, d->stream));
  PGO_CUDA(cudaMemcpyAsync(d_cov, cov, 9 * sizeof(double) * n, cudaMemcpyHostToDevice, d->stream));
  PGO_CUDA(cudaMemsetAsync(d->status.p, 0, sizeof(int), d->stream));
  label_star_edges<<<(n + 63) / 64, 64, 0, d->stream>>>(d->poses.p, gauge_vertex, n, d->scratch_i.p,
                                                        d_cov, d_meas, d_info, d->status.p);
  d->launches++;
  PGO_CUDA(cudaGetLastError());
  int status = 0;
  PGO_CUDA(cudaMemcpyAsync(&status, d->status.p, sizeof status, cudaMemcpyDeviceToHost, d->stream));
  PGO_CUDA(cudaMemcpyAsync(meas_out, d_meas, 3 * sizeof(double) * n, cudaMemcpyDeviceToHost, d->stream));
  PGO_CUDA(cudaMemcpyAsync(info_out, d_info, 9 * sizeof(double) * n, cudaMemcpyDeviceToHost, d->stream));
  PGO_CUDA(cudaStreamSynchronize(d->stream));
  if (status) {
    if (err) *err = "edge labelling: a covariance was not positive definite";
    return PGO_ERR_NUMERIC;
  }
  return PGO_OK;
}

}  // namespace pgo
