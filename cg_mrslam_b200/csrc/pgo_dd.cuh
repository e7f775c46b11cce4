// Domain-decomposed Gauss-Newton iteration (included by pgo_kernels.cu inside namespace pgo).
//
// One graph, N ranks (one per GPU). The structure analysis cuts the top log2(N) dissection levels
// into SHARED separator columns (ordered last) and gives every other column to one rank; no edge
// joins two ranks' interiors. One iteration is
//   gn_dd_local   (per rank)   zero, linearise the edges this rank owns, eliminate the rank's own
//                              columns (updates into shared columns accumulate in the rank's copy),
//                              forward-substitute the own rows, pack [rhs of the shared rows | chi2]
//                              behind the shared factor blocks;
//   ONE all-reduce (sum)       over [shared factor blocks | shared rhs | chi2] -- issued by the host
//                              (NCCL through torch.distributed) on the solver's stream;
//   gn_dd_shared  (replicated) eliminate the shared columns, finish the forward substitution,
//                              back-substitute shared then own columns, update own + shared poses.
// This is the "single all-reduce of the shared-separator Hessian rows" of the north star: the
// Schur complement of every rank's interior onto the separators is what gets summed.
//
// Edge ownership (who adds an edge's contribution to a SHARED diagonal block / rhs / chi2): the
// rank owning its interior endpoint; edges with no interior endpoint: edge index % world.

struct DDParams {
  const int* owner;  // per permuted column: rank or -1
  int rank, world;
  int first_shared;       // columns >= first_shared are shared
  int local_levels;       // phases 1..local_levels apply own-source updates
  int shared_min_level;
  const int* xfinal_ptr;
  const int* xfinal_cols;
  double* tail;           // right behind M: [3 * n_shared rhs | chi2 | pad]
  double* pose_x;         // [n_vertices][3] masked poses for the final exchange
};

__device__ __forceinline__ int edge_owner(const Params& P, const DDParams& D, int edge) {
  const int pi = P.vpos[P.edge_i[edge]], pj = P.vpos[P.edge_j[edge]];
  const int oi = pi >= 0 ? D.owner[pi] : -1, oj = pj >= 0 ? D.owner[pj] : -1;
  if (oi >= 0) return oi;
  if (oj >= 0) return oj;
  return edge % D.world;
}

__device__ void dd_linearise(const Params& P, const DDParams& D, double* scratch) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  double chi = 0.0;
  for (int p = tid; p < P.n; p += nthreads) {
    const int own = D.owner[p];
    if (own >= 0 && own != D.rank) continue;
    double diag[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
    for (int t = P.inc_ptr[p]; t < P.inc_ptr[p + 1]; ++t) {
      const Incidence inc = P.inc[t];
      const int edge = inc.edge_role >> 1, role = inc.edge_role & 1;
      if (edge_owner(P, D, edge) != D.rank) continue;
      EdgeLin L;
      linearise_edge(P, edge, &L);
      const double* jo = role ? L.jj : L.ji;
      const double* jx = role ? L.ji : L.jj;
      double A[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          A[3 * r + c] = jo[r] * L.om[c] + jo[3 + r] * L.om[3 + c] + jo[6 + r] * L.om[6 + c];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          diag[3 * r + c] += A[3 * r] * jo[c] + A[3 * r + 1] * jo[3 + c] + A[3 * r + 2] * jo[6 + c];
        b[r] -= A[3 * r] * L.e[0] + A[3 * r + 1] * L.e[1] + A[3 * r + 2] * L.e[2];
      }
      if (inc.pos >= 0) {
        double* dst = P.M + 9 * static_cast<size_t>(inc.pos);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            dst[3 * c + r] += A[3 * r] * jx[c] + A[3 * r + 1] * jx[3 + c] + A[3 * r + 2] * jx[6 + c];
      }
      if (role == 0 || P.vpos[P.edge_i[edge]] < 0) chi += edge_chi2(L);
    }
    double* d = P.M + 9 * static_cast<size_t>(P.col_ptr[p]);
#pragma unroll
    for (int i = 0; i < 9; ++i) d[i] = diag[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) P.rhs[3 * p + i] = b[i];
  }
  for (int t = tid; t < P.n_ff; t += nthreads) {
    if (edge_owner(P, D, P.ff_edges[t]) != D.rank) continue;
    EdgeLin L;
    linearise_edge(P, P.ff_edges[t], &L);
    chi += edge_chi2(L);
  }
  const double s = block_sum(chi, scratch);
  if (threadIdx.x == 0) P.chi2_partial[blockIdx.x] = s;
}

// mode 1: own-source updates (local stage); mode 2: shared-source updates (shared stage)
__device__ void dd_updates(const Params& P, const DDParams& D, int l, int mode) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  const int begin = P.phase_ptr[l], end = P.phase_ptr[l + 1];
  const int want = mode == 1 ? D.rank : -1;
  for (int i = begin + tid; i < end; i += nthreads) {
    const UpdateOp first = P.ops[i];
    if (i > begin && P.ops[i - 1].target == first.target) continue;
    const int target = first.target & ~kFinalFlag;
    const int tcol = P.col_of[target];
    const int towner = D.owner[tcol];
    // a target column owned by another rank can never receive one of our updates
    if (mode == 1 ? (towner >= 0 && towner != D.rank) : (towner >= 0)) continue;
    double acc[9];
    double* dst = P.M + 9 * static_cast<size_t>(target);
    load9(dst, acc);
    bool touched = false;
    int j = i;
    UpdateOp op = first;
    while (true) {
      if (D.owner[P.col_of[op.a]] == want) {
        double a[9], b[9], d[9], t[9];
        load9(P.M + 9 * static_cast<size_t>(op.a), a);
        load9(P.M + 9 * static_cast<size_t>(op.b), b);
        load9(P.Dinv + 9 * static_cast<size_t>(P.col_of[op.a]), d);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            t[3 * r + c] = a[3 * r] * d[c] + a[3 * r + 1] * d[3 + c] + a[3 * r + 2] * d[6 + c];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            acc[3 * r + c] -= t[3 * r] * b[3 * c] + t[3 * r + 1] * b[3 * c + 1] + t[3 * r + 2] * b[3 * c + 2];
        touched = true;
      }
      ++j;
      if (j >= end) break;
      op = P.ops[j];
      if (op.target != first.target) break;
    }
    if (touched) {
#pragma unroll
      for (int k = 0; k < 9; ++k) dst[k] = acc[k];
    }
    // the diagonal of a column of level l is complete after phase l -- for an own column in the
    // local stage, for a shared column in the shared stage (after the all-reduce)
    if ((first.target & kFinalFlag) && (mode == 1 ? towner == D.rank : towner < 0) && touched)
      finalise_diag(P, tcol, acc);
  }
}

__device__ void dd_leaves(const Params& P, const DDParams& D) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  for (int t = P.level_ptr[0] + tid; t < P.level_ptr[1]; t += nthreads) {
    const int col = P.level_cols[t];
    if (D.owner[col] != D.rank) continue;
    double m[9];
    load9(P.M + 9 * static_cast<size_t>(P.col_ptr[col]), m);
    finalise_diag(P, col, m);
  }
}

// shared columns of level l that no shared-source update completes: pivot them explicitly
__device__ void dd_explicit_finals(const Params& P, const DDParams& D, int l, bool solve_too) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  for (int t = D.xfinal_ptr[l] + tid; t < D.xfinal_ptr[l + 1]; t += nthreads) {
    const int col = D.xfinal_cols[t];
    if (!solve_too) {
      double m[9];
      load9(P.M + 9 * static_cast<size_t>(P.col_ptr[col]), m);
      finalise_diag(P, col, m);
    } else {
      finish_row(P, col, P.rhs + 3 * static_cast<size_t>(col), P.u + 3 * static_cast<size_t>(col));
    }
  }
}

__device__ void dd_forward_leaves(const Params& P, const DDParams& D) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  for (int t = P.level_ptr[0] + tid; t < P.level_ptr[1]; t += nthreads) {
    const int j = P.level_cols[t];
    if (D.owner[j] != D.rank) continue;
    finish_row(P, j, P.rhs + 3 * static_cast<size_t>(j), P.u + 3 * static_cast<size_t>(j));
  }
}

__device__ void dd_forward(const Params& P, const DDParams& D, int l, int mode) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  const int begin = P.fwd_ptr[l], end = P.fwd_ptr[l + 1];
  const int want = mode == 1 ? D.rank : -1;
  for (int i = begin + tid; i < end; i += nthreads) {
    const SolveOp first = P.fwd_ops[i];
    if (i > begin && P.fwd_ops[i - 1].row == first.row) continue;
    const int row = first.row & ~kFinalFlag;
    const int rowner = D.owner[row];
    if (mode == 1 ? (rowner >= 0 && rowner != D.rank) : (rowner >= 0)) continue;
    double* zz = P.rhs + 3 * static_cast<size_t>(row);
    double z0 = zz[0], z1 = zz[1], z2 = zz[2];
    bool touched = false;
    int j = i;
    SolveOp op = first;
    while (true) {
      const int k = P.col_of[op.pos];
      if (D.owner[k] == want) {
        const double* m = P.M + 9 * static_cast<size_t>(op.pos);
        const double* v = P.u + 3 * static_cast<size_t>(k);
        const double v0 = v[0], v1 = v[1], v2 = v[2];
        z0 -= m[0] * v0 + m[1] * v1 + m[2] * v2;
        z1 -= m[3] * v0 + m[4] * v1 + m[5] * v2;
        z2 -= m[6] * v0 + m[7] * v1 + m[8] * v2;
        touched = true;
      }
      ++j;
      if (j >= end) break;
      op = P.fwd_ops[j];
      if (op.row != first.row) break;
    }
    if (touched) {
      zz[0] = z0;
      zz[1] = z1;
      zz[2] = z2;
    }
    if ((first.row & kFinalFlag) && (mode == 1 ? rowner == D.rank : rowner < 0) && touched) {
      const double zf[3] = {z0, z1, z2};
      finish_row(P, row, zf, P.u + 3 * static_cast<size_t>(row));
    }
  }
}

// backward substitution of the columns of level l that are shared (mode 2) or own (mode 1)
__device__ void dd_backward(const Params& P, const DDParams& D, int l, int mode, double* scratch) {
  const int lane = threadIdx.x & 31;
  const int n_cols = P.level_ptr[l + 1] - P.level_ptr[l];
  const bool cta_mode = n_cols <= static_cast<int>(gridDim.x);
  const int group = cta_mode ? blockDim.x : 32;
  const int gid = cta_mode ? blockIdx.x : ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int n_groups = cta_mode ? gridDim.x : ((gridDim.x * blockDim.x) >> 5);
  const int rank_in = cta_mode ? threadIdx.x : lane;
  const int want = mode == 1 ? D.rank : -1;
  for (int w = gid; w < n_cols; w += n_groups) {
    const int j = P.level_cols[P.level_ptr[l] + w];
    if (D.owner[j] != want) continue;  // uniform over the group
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int t = P.col_ptr[j] + 1 + rank_in; t < P.col_ptr[j + 1]; t += group) {
      const double* m = P.M + 9 * static_cast<size_t>(t);
      const double* v = P.x + 3 * static_cast<size_t>(P.row_idx[t]);
      const double v0 = v[0], v1 = v[1], v2 = v[2];
      s0 += m[0] * v0 + m[3] * v1 + m[6] * v2;
      s1 += m[1] * v0 + m[4] * v1 + m[7] * v2;
      s2 += m[2] * v0 + m[5] * v1 + m[8] * v2;
    }
    for (int o = 16; o; o >>= 1) {
      s0 += __shfl_down_sync(0xffffffffu, s0, o);
      s1 += __shfl_down_sync(0xffffffffu, s1, o);
      s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    if (!cta_mode) {
      if (lane == 0) backward_store(P, j, 0, 0, s0, s1, s2);
    } else {
      const int warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
      if (lane == 0) {
        scratch[3 * warp] = s0;
        scratch[3 * warp + 1] = s1;
        scratch[3 * warp + 2] = s2;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
        for (int q = 0; q < n_warps; ++q) {
          t0 += scratch[3 * q];
          t1 += scratch[3 * q + 1];
          t2 += scratch[3 * q + 2];
        }
        backward_store(P, j, 0, 0, t0, t1, t2);
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kThreads) gn_dd_local(Params P, DDParams D) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double scratch[32];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  const int n_shared = P.n - D.first_shared;
  for (long long i = tid; i < P.nnzb * 9; i += nthreads) P.M[i] = 0.0;
  grid.sync();
  dd_linearise(P, D, scratch);
  grid.sync();
  if (blockIdx.x == 0) {
    double v = 0.0;
    for (int b = threadIdx.x; b < static_cast<int>(gridDim.x); b += blockDim.x) v += P.chi2_partial[b];
    const double s = block_sum(v, scratch);
    if (threadIdx.x == 0) D.tail[3 * static_cast<size_t>(n_shared)] = s;
  }
  dd_leaves(P, D);
  grid.sync();
  const int last = min(D.local_levels, P.n_levels - 1);
  for (int l = 1; l <= last; ++l) {
    dd_updates(P, D, l, 1);
    grid.sync();
  }
  dd_forward_leaves(P, D);
  grid.sync();
  for (int l = 1; l <= last; ++l) {
    dd_forward(P, D, l, 1);
    grid.sync();
  }
  // pack the shared rows' partial right-hand side behind the shared factor blocks
  for (int i = tid; i < 3 * n_shared; i += nthreads)
    D.tail[i] = P.rhs[3 * static_cast<size_t>(D.first_shared) + i];
  // a failed pivot on any rank must stop every rank: the flag rides in the exchange as well
  if (tid == 0)
    D.tail[3 * static_cast<size_t>(n_shared) + 1] = *reinterpret_cast<volatile int*>(P.status) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(kThreads) gn_dd_shared(Params P, DDParams D, int it) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double scratch[32];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  const int n_shared = P.n - D.first_shared;
  for (int i = tid; i < 3 * n_shared; i += nthreads)
    P.rhs[3 * static_cast<size_t>(D.first_shared) + i] = D.tail[i];
  if (D.tail[3 * static_cast<size_t>(n_shared) + 1] != 0.0) {  // some rank's interior failed
    if (tid == 0) atomicExch(P.status, 1);
    return;
  }
  if (tid == 0) P.chi2_out[it] = D.tail[3 * static_cast<size_t>(n_shared)];
  grid.sync();
  for (int l = D.shared_min_level; l < P.n_levels; ++l) {
    if (l >= 1) dd_updates(P, D, l, 2);
    dd_explicit_finals(P, D, l, false);
    grid.sync();
  }
  if (*reinterpret_cast<volatile int*>(P.status) != 0) return;
  for (int l = D.shared_min_level; l < P.n_levels; ++l) {
    if (l >= 1) dd_forward(P, D, l, 2);
    dd_explicit_finals(P, D, l, true);
    grid.sync();
  }
  for (int l = P.n_levels - 1; l >= D.shared_min_level; --l) {
    dd_backward(P, D, l, 2, scratch);
    grid.sync();
  }
  for (int l = min(D.local_levels, P.n_levels) - 1; l >= 0; --l) {
    dd_backward(P, D, l, 1, scratch);
    grid.sync();
  }
  for (int p = tid; p < P.n; p += nthreads) {
    const int own = D.owner[p];
    if (own >= 0 && own != D.rank) continue;
    const int v = P.perm_vertex[p];
    double* q = P.poses + 3 * static_cast<size_t>(v);
    q[0] += P.x[3 * p];
    q[1] += P.x[3 * p + 1];
    q[2] = normalize_theta(q[2] + P.x[3 * p + 2]);
  }
  if (tid == 0) P.status[1] = it + 1;
}

// Poses every rank is responsible for (own interior; shared + fixed ones from rank 0), zeros
// elsewhere: the sum over ranks is the full estimate vector.
__global__ void dd_mask_poses(Params P, DDParams D) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= P.n_vertices) return;
  const int p = P.vpos[v];
  const int own = p >= 0 ? D.owner[p] : -1;
  const bool mine = own >= 0 ? own == D.rank : D.rank == 0;
  for (int k = 0; k < 3; ++k) D.pose_x[3 * v + k] = mine ? P.poses[3 * v + k] : 0.0;
}
