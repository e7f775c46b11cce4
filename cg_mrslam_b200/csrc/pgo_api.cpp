// extern "C" surface of the pose-graph solver (include/pgo_solver.h): argument checking, active
// set / hessian-index bookkeeping, structure analysis (pgo_symbolic.cpp),
// spanning-tree initial guess (C10) and marshalling into the device back-end (pgo_device.h).
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <new>
#include <string>
#include <vector>

#include "../../include/pgo_solver.h"
#include "pgo_device.h"
#include "pgo_symbolic.h"

namespace {

thread_local std::string g_pgo_error;

int fail(int code, const std::string& msg) {
  g_pgo_error = msg;
  return code;
}

const double kPi = 3.14159265358979323846;

double normalize_theta(double t) {
  if (t >= -kPi && t < kPi) return t;
  const double two_pi = 2.0 * kPi;
  double w = t - two_pi * std::floor(t / two_pi);
  if (w >= kPi) w -= two_pi;
  return w;
}

struct Pose {
  double x, y, th;
};

// NCCL, bound at run time (libnccl.so.2: the copy already in the process -- a host program's, or
// PyTorch's -- else the system one), so that the library has no link-time dependency on it. Only
// the domain-decomposed solve (pgo_dd_*) needs it.
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;  // ncclUniqueId (nccl.h:37-38)
  int (*get_unique_id)(UniqueId*) = nullptr;
  int (*comm_init_rank)(void**, int, UniqueId, int) = nullptr;
  int (*all_reduce)(const void*, void*, size_t, int, int, void*, void*) = nullptr;
  int (*comm_destroy)(void*) = nullptr;
  const char* (*get_error_string)(int) = nullptr;
  bool ok = false;
};
const int kNcclDouble = 8, kNcclSum = 0;  // ncclFloat64, ncclSum (nccl.h:260,286)

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.get_unique_id = reinterpret_cast<int (*)(NcclApi::UniqueId*)>(dlsym(h, "ncclGetUniqueId"));
      api.comm_init_rank = reinterpret_cast<int (*)(void**, int, NcclApi::UniqueId, int)>(dlsym(h, "ncclCommInitRank"));
      api.all_reduce =
          reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, void*)>(dlsym(h, "ncclAllReduce"));
      api.comm_destroy = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclCommDestroy"));
      api.get_error_string = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
      api.ok = api.get_unique_id && api.comm_init_rank && api.all_reduce && api.comm_destroy;
    }
  }
  return api;
}

Pose se2_inv(const Pose& a) {  // C1
  Pose r;
  r.th = normalize_theta(-a.th);
  const double c = std::cos(r.th), s = std::sin(r.th);
  r.x = c * (-a.x) - s * (-a.y);
  r.y = s * (-a.x) + c * (-a.y);
  return r;
}

Pose se2_mul(const Pose& a, const Pose& b) {  // C1
  const double c = std::cos(a.th), s = std::sin(a.th);
  Pose r;
  r.x = a.x + (c * b.x - s * b.y);
  r.y = a.y + (s * b.x + c * b.y);
  r.th = normalize_theta(a.th + b.th);
  return r;
}

}  // namespace

struct pgo_solver {
  pgo::DeviceSolver* dev = nullptr;
  int n_vertices = 0, n_edges = 0;
  std::vector<int> edge_i, edge_j;
  std::vector<uint8_t> fixed;
  std::vector<int> hidx;   // per vertex, -1 if fixed
  pgo::Symbolic sym;
  pgo::GraphTables tables;
  std::vector<double> meas;  // host copy for the initial guess
  bool have_graph = false, have_values = false;
  double last_ms = 0.0;
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;   // ncclComm_t of the domain decomposition (pgo_dd_set_comm / pgo_dd_comm_init)
  bool own_comm = false;
  int analysed_rank = -1, analysed_world = -1, analysed_batch = -1;  // what the stored analysis was made for
  long long structure_hits = 0;  // pgo_set_graph calls answered from the stored analysis
};

extern "C" {

const char* pgo_last_error(void) { return g_pgo_error.c_str(); }

int pgo_create(pgo_solver** out, int device, void* stream) {
  if (!out) return fail(PGO_ERR_ARG, "out is null");
  *out = nullptr;
  pgo_solver* s = new (std::nothrow) pgo_solver();
  if (!s) return fail(PGO_ERR_ALLOC, "out of host memory");
  std::string err;
  const int rc = pgo::dev_create(&s->dev, device, stream, &err);
  if (rc) {
    delete s;
    return fail(rc, err);
  }
  *out = s;
  return PGO_OK;
}

void pgo_destroy(pgo_solver* s) {
  if (!s) return;
  if (s->nccl_comm && s->own_comm && nccl().ok) nccl().comm_destroy(s->nccl_comm);
  pgo::dev_destroy(s->dev);
  delete s;
}

int pgo_set_graph(pgo_solver* s, int n_vertices, int n_edges, const int32_t* edge_i,
                  const int32_t* edge_j, const uint8_t* fixed) {
  if (!s || n_vertices <= 0 || n_edges < 0 || (n_edges && (!edge_i || !edge_j)) || !fixed)
    return fail(PGO_ERR_ARG, "bad argument");
  // The keyframe loop re-initialises the optimiser before every optimize() call
  // (graph_slam.cpp:388-393,561-565): when nothing changed since the last call -- same vertices,
  // same edges, same fixed flags -- the analysis, the device tables and the captured iteration
  // graph are all still valid.
  if (s->have_graph && n_vertices == s->n_vertices && n_edges == s->n_edges && s->analysed_world == s->world &&
      s->analysed_rank == s->rank && s->analysed_batch == pgo::dev_batch(s->dev) &&
      std::equal(fixed, fixed + n_vertices, s->fixed.begin()) &&
      std::equal(edge_i, edge_i + n_edges, s->edge_i.begin()) &&
      std::equal(edge_j, edge_j + n_edges, s->edge_j.begin())) {
    ++s->structure_hits;
    return PGO_OK;
  }
  s->have_graph = s->have_values = false;
  s->n_vertices = n_vertices;
  s->n_edges = n_edges;
  s->edge_i.assign(edge_i, edge_i + n_edges);
  s->edge_j.assign(edge_j, edge_j + n_edges);
  s->fixed.assign(fixed, fixed + n_vertices);
  // hessian indices: non-fixed vertices 0..n-1 in vertex order (C6)
  s->hidx.assign(n_vertices, -1);
  int n = 0;
  for (int v = 0; v < n_vertices; ++v)
    if (!fixed[v]) s->hidx[v] = n++;
  std::vector<std::pair<int, int> > free_edges;
  free_edges.reserve(n_edges);
  for (int e = 0; e < n_edges; ++e) {
    const int i = edge_i[e], j = edge_j[e];
    if (i < 0 || i >= n_vertices || j < 0 || j >= n_vertices)
      return fail(PGO_ERR_ARG, "edge endpoint out of range");
    if (i == j) return fail(PGO_ERR_ARG, "EdgeSE2 with identical endpoints");
    if (s->hidx[i] >= 0 && s->hidx[j] >= 0)
      free_edges.push_back(std::make_pair(s->hidx[i], s->hidx[j]));
  }
  std::string err;
  if (s->world > 1 && pgo::dev_batch(s->dev) != 1)
    return fail(PGO_ERR_ARG, "a domain-decomposed solver holds one graph instance (pgo_set_batch(1))");
  if (!pgo::analyse(n, free_edges, 0, s->world, &s->sym, &err)) return fail(PGO_ERR_CAPACITY, err);
  const pgo::Symbolic& S = s->sym;

  pgo::GraphTables& G = s->tables;
  G = pgo::GraphTables();
  G.n_vertices = n_vertices;
  G.n_edges = n_edges;
  G.edge_i = s->edge_i;
  G.edge_j = s->edge_j;
  G.vpos.assign(n_vertices, -1);
  for (int v = 0; v < n_vertices; ++v)
    if (s->hidx[v] >= 0) G.vpos[v] = S.iperm[s->hidx[v]];
  G.inc_ptr.assign(n + 1, 0);
  for (int e = 0; e < n_edges; ++e) {
    const int pi = G.vpos[edge_i[e]], pj = G.vpos[edge_j[e]];
    if (pi >= 0) ++G.inc_ptr[pi + 1];
    if (pj >= 0) ++G.inc_ptr[pj + 1];
    if (pi < 0 && pj < 0) G.ff_edges.push_back(e);
  }
  for (int p = 0; p < n; ++p) G.inc_ptr[p + 1] += G.inc_ptr[p];
  G.inc.resize(G.inc_ptr[n]);
  std::vector<int> cur(G.inc_ptr.begin(), G.inc_ptr.end() - 1);
  std::vector<char> seen_block(S.nnzb, 0);
  int64_t off_blocks = 0;
  for (int e = 0; e < n_edges; ++e) {  // ascending edge index = g2o's accumulation order (C5)
    const int pi = G.vpos[edge_i[e]], pj = G.vpos[edge_j[e]];
    int pos = -1;
    if (pi >= 0 && pj >= 0) {
      const int c = std::min(pi, pj), r = std::max(pi, pj);
      const int* b = S.row_idx.data() + S.col_ptr[c] + 1;
      const int* en = S.row_idx.data() + S.col_ptr[c + 1];
      const int* it = std::lower_bound(b, en, r);
      if (it == en || *it != r) return fail(PGO_ERR_CAPACITY, "internal: H block missing in factor");
      pos = static_cast<int>(it - S.row_idx.data());
      if (!seen_block[pos]) {
        seen_block[pos] = 1;
        ++off_blocks;
      }
    }
    if (pi >= 0) {
      pgo::Incidence x = {2 * e, (pj >= 0 && pi < pj) ? pos : -1};
      G.inc[cur[pi]++] = x;
    }
    if (pj >= 0) {
      pgo::Incidence x = {2 * e + 1, (pi >= 0 && pj < pi) ? pos : -1};
      G.inc[cur[pj]++] = x;
    }
  }
  G.hessian_blocks = n + off_blocks;
  G.rank = s->rank;
  const int rc = pgo::dev_set_structure(s->dev, S, G, &err);
  if (rc) return fail(rc, err);
  s->have_graph = true;
  s->analysed_rank = s->rank;
  s->analysed_world = s->world;
  s->analysed_batch = pgo::dev_batch(s->dev);
  return PGO_OK;
}

int pgo_set_batch(pgo_solver* s, int batch) {
  if (!s) return fail(PGO_ERR_ARG, "null solver");
  std::string err;
  const int rc = pgo::dev_set_batch(s->dev, batch, &err);
  if (rc) return fail(rc, err);
  s->have_graph = s->have_values = false;
  return PGO_OK;
}

static int upload_impl(pgo_solver* s, int inst, const double* poses, const double* meas,
                       const double* info6) {
  if (!s || !s->have_graph || !poses || (s->n_edges && (!meas || !info6)))
    return fail(PGO_ERR_ARG, "bad argument (pgo_set_graph first)");
  std::string err;
  const int rc = pgo::dev_upload(s->dev, inst, poses, meas, info6, &err);
  if (rc) return fail(rc, err);
  if (inst <= 0) {  // the host copy of the measurements (initial guess) follows instance 0
    s->meas.assign(meas, meas + 3 * static_cast<size_t>(s->n_edges));
    s->have_values = true;
  }
  return PGO_OK;
}

int pgo_upload(pgo_solver* s, const double* poses, const double* meas, const double* info6) {
  return upload_impl(s, -1, poses, meas, info6);
}

int pgo_upload_instance(pgo_solver* s, int inst, const double* poses, const double* meas,
                        const double* info6) {
  if (inst < 0) return fail(PGO_ERR_ARG, "no such instance");
  return upload_impl(s, inst, poses, meas, info6);
}

int pgo_set_poses(pgo_solver* s, const double* poses) {
  if (!s || !s->have_values || !poses) return fail(PGO_ERR_ARG, "bad argument (pgo_upload first)");
  std::string err;
  const int rc = pgo::dev_set_poses(s->dev, -1, poses, &err);
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_get_poses(pgo_solver* s, double* poses) { return pgo_get_poses_instance(s, 0, poses); }

int pgo_get_poses_instance(pgo_solver* s, int inst, double* poses) {
  if (!s || !s->have_values || !poses) return fail(PGO_ERR_ARG, "bad argument (pgo_upload first)");
  std::string err;
  const int rc = pgo::dev_get_poses(s->dev, inst, poses, &err);
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_iterate_batch(pgo_solver* s, int n_iters, double* chi2_out, int* iters_done) {
  if (!s || !s->have_values || n_iters < 0 || !iters_done)
    return fail(PGO_ERR_ARG, "bad argument (pgo_upload first)");
  std::string err;
  float ms = 0.f;
  const int rc = pgo::dev_iterate(s->dev, n_iters, chi2_out, iters_done, &ms, &err);
  s->last_ms = ms;
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_iterate(pgo_solver* s, int n_iters, double* poses_out, double* chi2_out, int* iters_done) {
  if (!s || !s->have_values || n_iters < 0) return fail(PGO_ERR_ARG, "bad argument (pgo_upload first)");
  std::string err;
  const int batch = pgo::dev_batch(s->dev);
  std::vector<int> done(batch, 0);
  std::vector<double> chi2(chi2_out ? static_cast<size_t>(batch) * std::max(n_iters, 1) : 0);
  float ms = 0.f;
  int rc = pgo::dev_iterate(s->dev, n_iters, chi2_out ? chi2.data() : nullptr, done.data(), &ms, &err);
  s->last_ms = ms;
  if (iters_done) *iters_done = done[0];
  if (chi2_out) std::copy(chi2.begin(), chi2.begin() + n_iters, chi2_out);  // instance 0
  if (rc && rc != PGO_ERR_NUMERIC) return fail(rc, err);
  if (poses_out) {
    std::string err2;
    const int rc2 = pgo::dev_get_poses(s->dev, 0, poses_out, &err2);
    if (rc2) return fail(rc2, err2);
  }
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_chi2(pgo_solver* s, double* chi2) {
  if (!s || !s->have_values || !chi2) return fail(PGO_ERR_ARG, "bad argument (pgo_upload first)");
  std::string err;
  const int rc = pgo::dev_chi2(s->dev, chi2, &err);
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_marginals(pgo_solver* s, int n_blocks, const int32_t* vertex_r, const int32_t* vertex_c,
                  double* cov_out) {
  if (!s || !s->have_values || n_blocks < 0 || (n_blocks && (!vertex_r || !vertex_c || !cov_out)))
    return fail(PGO_ERR_ARG, "bad argument");
  std::vector<int> rp(n_blocks), cp(n_blocks);
  for (int k = 0; k < n_blocks; ++k) {
    const int r = vertex_r[k], c = vertex_c[k];
    if (r < 0 || r >= s->n_vertices || c < 0 || c >= s->n_vertices || s->hidx[r] < 0 || s->hidx[c] < 0)
      return fail(PGO_ERR_ARG, "marginal requested for a fixed or unknown vertex");
    rp[k] = s->tables.vpos[r];
    cp[k] = s->tables.vpos[c];
  }
  std::string err;
  const int rc = pgo::dev_marginals(s->dev, n_blocks, cp.data(), rp.data(), cov_out, &err);
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_initial_guess(pgo_solver* s) {
  if (!s || !s->have_values) return fail(PGO_ERR_ARG, "bad argument (pgo_upload first)");
  const int nv = s->n_vertices, ne = s->n_edges;
  std::vector<double> poses(3 * static_cast<size_t>(nv));
  std::string err;
  int rc = pgo::dev_get_poses(s->dev, 0, poses.data(), &err);
  if (rc) return fail(rc, err);
  // C10 with a tie rule that does not depend on the order in which the caller lists the edges (g2o's
  // depends on pointer values): breadth-first by levels from the fixed vertices; the vertices of a
  // level are expanded in ascending index, the edges of a vertex in ascending (other end, direction,
  // measurement); the first edge to reach a vertex sets it.
  std::vector<int> ptr(nv + 1, 0);
  for (int e = 0; e < ne; ++e) {
    ++ptr[s->edge_i[e] + 1];
    ++ptr[s->edge_j[e] + 1];
  }
  for (int v = 0; v < nv; ++v) ptr[v + 1] += ptr[v];
  std::vector<int> adj(ptr[nv]), cur(ptr.begin(), ptr.end() - 1);
  for (int e = 0; e < ne; ++e) {
    adj[cur[s->edge_i[e]]++] = 2 * e;      // forward: this vertex is Xi
    adj[cur[s->edge_j[e]]++] = 2 * e + 1;  // backward: this vertex is Xj
  }
  const double* meas = s->meas.data();
  const int32_t* ei = s->edge_i.data();
  const int32_t* ej = s->edge_j.data();
  auto before = [meas, ei, ej](int a, int b) {
    const int ea = a >> 1, eb = b >> 1;
    const int wa = (a & 1) ? ei[ea] : ej[ea], wb = (b & 1) ? ei[eb] : ej[eb];
    if (wa != wb) return wa < wb;
    if ((a & 1) != (b & 1)) return (a & 1) < (b & 1);
    for (int k = 0; k < 3; ++k)
      if (meas[3 * ea + k] != meas[3 * eb + k]) return meas[3 * ea + k] < meas[3 * eb + k];
    return false;
  };
  for (int v = 0; v < nv; ++v) std::stable_sort(adj.begin() + ptr[v], adj.begin() + ptr[v + 1], before);
  std::vector<char> seen(nv, 0);
  std::vector<int> level, next;
  for (int v = 0; v < nv; ++v)
    if (s->fixed[v]) {
      seen[v] = 1;
      level.push_back(v);
    }
  while (!level.empty()) {
    next.clear();
    for (size_t q = 0; q < level.size(); ++q) {
      const int v = level[q];
      const Pose xv = {poses[3 * v], poses[3 * v + 1], poses[3 * v + 2]};
      for (int t = ptr[v]; t < ptr[v + 1]; ++t) {
        const int e = adj[t] >> 1, forward = !(adj[t] & 1);
        const int w = forward ? s->edge_j[e] : s->edge_i[e];
        if (seen[w]) continue;
        const Pose z = {s->meas[3 * e], s->meas[3 * e + 1], s->meas[3 * e + 2]};
        const Pose xw = forward ? se2_mul(xv, z) : se2_mul(xv, se2_inv(z));  // EdgeSE2::initialEstimate
        poses[3 * w] = xw.x;
        poses[3 * w + 1] = xw.y;
        poses[3 * w + 2] = xw.th;
        seen[w] = 1;
        next.push_back(w);
      }
    }
    std::sort(next.begin(), next.end());
    level.swap(next);
  }
  rc = pgo::dev_set_poses(s->dev, 0, poses.data(), &err);
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_label_star_edges(pgo_solver* s, int gauge, int n, const int32_t* v, double* meas_out,
                         double* info_out) {
  if (!s || !s->have_values || n < 0 || (n && (!v || !meas_out || !info_out)) || gauge < 0 ||
      gauge >= s->n_vertices)
    return fail(PGO_ERR_ARG, "bad argument");
  if (n == 0) return PGO_OK;
  std::vector<double> cov(9 * static_cast<size_t>(n));
  int rc = pgo_marginals(s, n, v, v, cov.data());
  if (rc) return rc;
  std::string err;
  rc = pgo::dev_label_star(s->dev, gauge, n, v, cov.data(), meas_out, info_out, &err);
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_set_partition(pgo_solver* s, int rank, int world) {
  if (!s || world < 1 || (world & (world - 1)) != 0 || rank < 0 || rank >= world)
    return fail(PGO_ERR_ARG, "world must be a power of two and 0 <= rank < world");
  s->rank = rank;
  s->world = world;
  s->have_graph = s->have_values = false;  // takes effect at the next pgo_set_graph
  return PGO_OK;
}

int pgo_dd_begin(pgo_solver* s, int n_iters) {
  if (!s || !s->have_values || n_iters < 0) return fail(PGO_ERR_ARG, "bad argument (pgo_upload first)");
  std::string err;
  const int rc = pgo::dev_dd_begin(s->dev, n_iters, &err);
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_dd_local(pgo_solver* s) {
  if (!s || !s->have_values) return fail(PGO_ERR_ARG, "bad argument (pgo_upload first)");
  std::string err;
  const int rc = pgo::dev_dd_local(s->dev, &err);
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_dd_exchange_buffer(pgo_solver* s, void** dev_ptr, int64_t* n_doubles) {
  if (!s || !s->have_graph || !dev_ptr || !n_doubles) return fail(PGO_ERR_ARG, "bad argument");
  long long n = 0;
  pgo::dev_dd_exchange(s->dev, dev_ptr, &n);
  *n_doubles = n;
  return PGO_OK;
}

int pgo_dd_shared(pgo_solver* s) {
  if (!s || !s->have_values) return fail(PGO_ERR_ARG, "bad argument (pgo_upload first)");
  std::string err;
  const int rc = pgo::dev_dd_shared(s->dev, &err);
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_dd_end(pgo_solver* s, int n_iters, double* chi2_out, int* iters_done) {
  if (!s || !s->have_values) return fail(PGO_ERR_ARG, "bad argument (pgo_upload first)");
  std::string err;
  int done = 0;
  float ms = 0.f;
  const int rc = pgo::dev_dd_end(s->dev, n_iters, chi2_out, &done, &ms, &err);
  s->last_ms = ms;
  if (iters_done) *iters_done = done;
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_dd_pose_exchange(pgo_solver* s, void** dev_ptr, int64_t* n_doubles) {
  if (!s || !s->have_values || !dev_ptr || !n_doubles) return fail(PGO_ERR_ARG, "bad argument");
  std::string err;
  long long n = 0;
  const int rc = pgo::dev_dd_pose_exchange(s->dev, dev_ptr, &n, &err);
  *n_doubles = n;
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_dd_pose_commit(pgo_solver* s) {
  if (!s || !s->have_values) return fail(PGO_ERR_ARG, "bad argument");
  std::string err;
  const int rc = pgo::dev_dd_pose_commit(s->dev, &err);
  return rc ? fail(rc, err) : PGO_OK;
}

int pgo_dd_unique_id(void* out128) {
  if (!out128) return fail(PGO_ERR_ARG, "null argument");
  if (!nccl().ok) return fail(PGO_ERR_CUDA, "libnccl.so.2 could not be loaded");
  const int rc = nccl().get_unique_id(static_cast<NcclApi::UniqueId*>(out128));
  return rc ? fail(PGO_ERR_CUDA, std::string("ncclGetUniqueId: ") + nccl().get_error_string(rc)) : PGO_OK;
}

int pgo_dd_comm_init(pgo_solver* s, const void* unique_id128, int rank, int world) {
  if (!s || !unique_id128 || world < 1 || rank < 0 || rank >= world) return fail(PGO_ERR_ARG, "bad argument");
  if (!nccl().ok) return fail(PGO_ERR_CUDA, "libnccl.so.2 could not be loaded");
  if (s->nccl_comm && s->own_comm) nccl().comm_destroy(s->nccl_comm);
  s->nccl_comm = nullptr;
  NcclApi::UniqueId id;
  std::memcpy(&id, unique_id128, sizeof id);
  const int rc = nccl().comm_init_rank(&s->nccl_comm, world, id, rank);
  if (rc) return fail(PGO_ERR_CUDA, std::string("ncclCommInitRank: ") + nccl().get_error_string(rc));
  s->own_comm = true;
  return pgo_set_partition(s, rank, world);
}

int pgo_dd_set_comm(pgo_solver* s, void* nccl_comm, int rank, int world) {
  if (!s || !nccl_comm) return fail(PGO_ERR_ARG, "bad argument");
  if (!nccl().ok) return fail(PGO_ERR_CUDA, "libnccl.so.2 could not be loaded");
  if (s->nccl_comm && s->own_comm) nccl().comm_destroy(s->nccl_comm);
  s->nccl_comm = nccl_comm;
  s->own_comm = false;
  return pgo_set_partition(s, rank, world);
}

int pgo_dd_iterate(pgo_solver* s, int n_iters, double* chi2_out, int* iters_done) {
  if (!s || !s->have_values || n_iters < 0) return fail(PGO_ERR_ARG, "bad argument (pgo_upload first)");
  if (!s->nccl_comm) return fail(PGO_ERR_ARG, "no communicator (pgo_dd_comm_init / pgo_dd_set_comm first)");
  std::string err;
  void* stream = pgo::dev_stream(s->dev);
  int rc = pgo::dev_dd_begin(s->dev, n_iters, &err);
  if (rc) return fail(rc, err);
  void* ptr = nullptr;
  long long n = 0;
  for (int it = 0; it < n_iters; ++it) {
    rc = pgo::dev_dd_local(s->dev, &err);
    if (rc) return fail(rc, err);
    pgo::dev_dd_exchange(s->dev, &ptr, &n);
    int nrc = nccl().all_reduce(ptr, ptr, static_cast<size_t>(n), kNcclDouble, kNcclSum, s->nccl_comm, stream);
    if (nrc) return fail(PGO_ERR_CUDA, std::string("ncclAllReduce: ") + nccl().get_error_string(nrc));
    rc = pgo::dev_dd_shared(s->dev, &err);
    if (rc) return fail(rc, err);
  }
  int done = 0;
  float ms = 0.f;
  const int rc_end = pgo::dev_dd_end(s->dev, n_iters, chi2_out, &done, &ms, &err);
  s->last_ms = ms;
  if (iters_done) *iters_done = done;
  if (rc_end && rc_end != PGO_ERR_NUMERIC) return fail(rc_end, err);
  // every rank ends with every estimate: owners contribute theirs, the rest zeros
  std::string err2;
  rc = pgo::dev_dd_pose_exchange(s->dev, &ptr, &n, &err2);
  if (rc) return fail(rc, err2);
  const int nrc = nccl().all_reduce(ptr, ptr, static_cast<size_t>(n), kNcclDouble, kNcclSum, s->nccl_comm, stream);
  if (nrc) return fail(PGO_ERR_CUDA, std::string("ncclAllReduce: ") + nccl().get_error_string(nrc));
  rc = pgo::dev_dd_pose_commit(s->dev, &err2);
  if (rc) return fail(rc, err2);
  return rc_end ? fail(rc_end, err) : PGO_OK;
}

int pgo_analyse_partition(int n_vertices, int n_edges, const int32_t* edge_i, const int32_t* edge_j,
                          const uint8_t* fixed, int world, int32_t* vertex_owner, int64_t* stats) {
  if (n_vertices <= 0 || n_edges < 0 || (n_edges && (!edge_i || !edge_j)) || !fixed || !vertex_owner)
    return fail(PGO_ERR_ARG, "bad argument");
  std::vector<int> hidx(n_vertices, -1);
  int n = 0;
  for (int v = 0; v < n_vertices; ++v)
    if (!fixed[v]) hidx[v] = n++;
  std::vector<std::pair<int, int> > free_edges;
  for (int e = 0; e < n_edges; ++e) {
    const int i = edge_i[e], j = edge_j[e];
    if (i < 0 || i >= n_vertices || j < 0 || j >= n_vertices || i == j)
      return fail(PGO_ERR_ARG, "bad edge");
    if (hidx[i] >= 0 && hidx[j] >= 0) free_edges.push_back(std::make_pair(hidx[i], hidx[j]));
  }
  pgo::Symbolic S;
  std::string err;
  if (!pgo::analyse(n, free_edges, 0, world, &S, &err)) return fail(PGO_ERR_CAPACITY, err);
  for (int v = 0; v < n_vertices; ++v)
    vertex_owner[v] = hidx[v] < 0 ? -2 : S.owner[S.iperm[hidx[v]]];
  if (stats) {
    stats[0] = S.n - S.first_shared;
    stats[1] = S.nnzb - (S.n > S.first_shared ? S.col_ptr[S.first_shared] : S.nnzb);
    for (int r = 0; r <= world; ++r) stats[2 + r] = 0;
    for (int p = 0; p < S.n; ++p) {
      const int64_t m = S.col_ptr[p + 1] - S.col_ptr[p] - 1;
      stats[2 + (S.owner[p] < 0 ? world : S.owner[p])] += m * (m + 1) / 2;
    }
  }
  return PGO_OK;
}

int pgo_get_stats(const pgo_solver* s, pgo_stats* out) {
  if (!s || !out) return fail(PGO_ERR_ARG, "null argument");
  std::memset(out, 0, sizeof *out);
  out->n_vertices = s->n_vertices;
  out->n_edges = s->n_edges;
  out->n_free = s->sym.n;
  out->n_levels = s->sym.n_levels;
  out->factor_blocks = s->sym.nnzb;
  out->update_ops = s->sym.n_ops;
  out->hessian_blocks = s->tables.hessian_blocks;
  out->analyse_seconds = s->sym.analyse_seconds;
  out->last_iterate_ms = s->last_ms;
  out->kernel_launches = static_cast<int64_t>(pgo::dev_launches(s->dev));
  for (int k = 0; k < 5; ++k) out->stage_ms[k] = pgo::dev_stage_ms(s->dev)[k];
  out->factor_flops = s->sym.sn.flops;
  out->n_supernodes = s->sym.sn.n_super;
  out->n_panels = s->sym.sn.n_panels;
  out->n_panel_levels = s->sym.sn.n_plevels;
  out->n_supernode_levels = s->sym.sn.n_slevels;
  out->batch = pgo::dev_batch(s->dev);
  return PGO_OK;
}

void* pgo_stream(const pgo_solver* s) { return s ? pgo::dev_stream(s->dev) : nullptr; }

}  // extern "C"
