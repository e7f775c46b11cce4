// Domain-decomposed Gauss-Newton iteration (included by pgo_kernels.cu inside namespace pgo).
//
// One graph, N ranks (one per GPU). The structure analysis cuts the top log2(N) dissection levels
// into SHARED separator columns (ordered last) and gives every other column to one rank; no edge
// joins two ranks' interiors. One iteration is
//   local stage   (per rank)   zero, linearise the edges this rank owns, factorise + forward-
//                              substitute the rank's own panels with the supernodal kernels (their
//                              outer products into shared columns accumulate in the rank's copy),
//                              pack [rhs of the shared rows | chi2 | failure flag] behind the
//                              shared factor blocks;
//   ONE all-reduce (sum)       over [shared factor blocks | shared rhs | chi2] -- issued by the host
//                              (NCCL through torch.distributed) on the solver's stream;
//   shared stage  (replicated) factorise the shared panels, back-substitute shared then own
//                              supernodes, update own + shared poses.
// Both stages are CUDA graphs of the per-phase kernels of pgo_kernels.cu, restricted to the task
// lists of this rank's panels / of the shared panels (Supernodal::lists(owner)).
// This is the "single all-reduce of the shared-separator Hessian rows" of the north star: the
// Schur complement of every rank's interior onto the separators is what gets summed.
//
// Edge ownership (who adds an edge's contribution to a SHARED diagonal block / rhs / chi2): the
// rank owning its interior endpoint; edges with no interior endpoint: edge index % world.

struct DDParams {
  const int* owner;  // per permuted column: rank or -1
  int rank, world;
  int first_shared;       // columns >= first_shared are shared
  double* tail;           // right behind M: [3 * n_shared rhs | chi2 | failure flag | pad]
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

__global__ void __launch_bounds__(kThreads) gn_dd_linearise(Params P, DDParams D) {
  __shared__ double scratch[32];
  dd_linearise(P, D, scratch);
}

// chi2 of this rank's edges, the partial right-hand side of the shared rows and the failure flag
// go behind the shared factor blocks: they ride in the same all-reduce
__global__ void gn_dd_pack(Params P, DDParams D, int n_partials) {
  __shared__ double scratch[32];
  const int n_shared = P.n - D.first_shared;
  for (int i = threadIdx.x; i < 3 * n_shared; i += blockDim.x)
    D.tail[i] = P.rhs[3 * static_cast<size_t>(D.first_shared) + i];
  double v = 0.0;
  for (int b = threadIdx.x; b < n_partials; b += blockDim.x) v += P.chi2_partial[b];
  const double s = block_sum(v, scratch);
  if (threadIdx.x == 0) {
    D.tail[3 * static_cast<size_t>(n_shared)] = s;
    // a failed pivot on any rank must stop every rank
    D.tail[3 * static_cast<size_t>(n_shared) + 1] = *reinterpret_cast<volatile int*>(P.status) ? 1.0 : 0.0;
  }
}

__global__ void gn_dd_unpack(Params P, DDParams D) {
  const int n_shared = P.n - D.first_shared;
  if (D.tail[3 * static_cast<size_t>(n_shared) + 1] != 0.0) {  // some rank's interior failed
    if (threadIdx.x == 0) atomicExch(P.status, 1);
    return;
  }
  for (int i = threadIdx.x; i < 3 * n_shared; i += blockDim.x)
    P.rhs[3 * static_cast<size_t>(D.first_shared) + i] = D.tail[i];
  if (threadIdx.x == 0) P.chi2_out[P.status[1]] = D.tail[3 * static_cast<size_t>(n_shared)];
}

// oplus on the vertices this rank holds: its own interior and the (replicated) separators
__global__ void gn_dd_update(Params P, DDParams D) {
  if (*reinterpret_cast<volatile int*>(P.status) != 0) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.n) return;
  const int own = D.owner[p];
  if (own >= 0 && own != D.rank) return;
  const int v = P.perm_vertex[p];
  double* q = P.poses + 3 * static_cast<size_t>(v);
  q[0] += P.x[3 * p];
  q[1] += P.x[3 * p + 1];
  q[2] = normalize_theta(q[2] + P.x[3 * p + 2]);
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
