// Device back-end interface of the pose-graph solver. Implemented by pgo_kernels.cu (CUDA,
// sm_100a); there is no other implementation.
#ifndef CGM_PGO_DEVICE_H
#define CGM_PGO_DEVICE_H

#include <cstdint>
#include <string>
#include <vector>

#include "pgo_symbolic.h"

namespace pgo {

struct DeviceSolver;  // opaque

// Host-built description of the linearisation: incidences grouped by permuted free vertex.
struct Incidence {
  int edge_role;  // edge index * 2 + role (0: this vertex is the edge's i, 1: its j)
  int pos;        // factor position of the off-diagonal block this visit accumulates, or -1
};

struct GraphTables {
  int n_vertices = 0, n_edges = 0;
  std::vector<int> edge_i, edge_j;
  std::vector<int> vpos;          // per vertex: permuted position p, or -1 if fixed
  std::vector<int> inc_ptr;       // n + 1, by permuted position
  std::vector<Incidence> inc;
  std::vector<int> ff_edges;      // edges between two fixed vertices (chi2 only)
  int64_t hessian_blocks = 0;
  int rank = 0;  // this process' rank in the domain decomposition (Symbolic::world ranks)
};

int dev_create(DeviceSolver** out, int device, void* stream, std::string* err);
void dev_destroy(DeviceSolver* d);
void* dev_stream(const DeviceSolver* d);
uint64_t dev_launches(const DeviceSolver* d);
// Device time of the stages of the last GN iteration: linearise (incl. zero fill and leaf pivots),
// factor updates, forward solve, backward solve, update.
const double* dev_stage_ms(const DeviceSolver* d);

int dev_set_structure(DeviceSolver* d, const Symbolic& S, const GraphTables& G, std::string* err);
// A solver can hold a batch of problem instances with the same structure (different estimates,
// measurements, information matrices); every kernel of the iteration graph then runs all of them
// (blockIdx.y = instance). Set before dev_set_structure. inst = -1 addresses every instance.
int dev_set_batch(DeviceSolver* d, int batch, std::string* err);
int dev_batch(const DeviceSolver* d);
int dev_upload(DeviceSolver* d, int inst, const double* poses, const double* meas,
               const double* info6, std::string* err);
int dev_set_poses(DeviceSolver* d, int inst, const double* poses, std::string* err);
int dev_get_poses(DeviceSolver* d, int inst, double* poses, std::string* err);

// n_iters Gauss-Newton iterations of every instance (one CUDA graph launch per iteration).
// chi2_out[batch][n_iters] (host, may be null), iters_done[batch]: < n_iters iff a diagonal block
// of that instance was not positive definite. *ms = device time.
int dev_iterate(DeviceSolver* d, int n_iters, double* chi2_out, int* iters_done, float* ms,
                std::string* err);
// Domain-decomposed iteration (Symbolic::world > 1), all asynchronous on the solver's stream:
//   begin; per iteration { local; <all-reduce(sum) of the exchange buffer by the caller>; shared };
//   end (synchronises); then pose_exchange + <all-reduce> + pose_commit to assemble all estimates.
int dev_dd_begin(DeviceSolver* d, int n_iters, std::string* err);
int dev_dd_local(DeviceSolver* d, std::string* err);
void dev_dd_exchange(DeviceSolver* d, void** ptr, long long* n_doubles);
int dev_dd_shared(DeviceSolver* d, std::string* err);
int dev_dd_end(DeviceSolver* d, int n_iters, double* chi2_out, int* iters_done, float* ms,
               std::string* err);
int dev_dd_pose_exchange(DeviceSolver* d, void** ptr, long long* n_doubles, std::string* err);
int dev_dd_pose_commit(DeviceSolver* d, std::string* err);
int dev_chi2(DeviceSolver* d, double* chi2, std::string* err);
// Solve H X = E for nrhs unit-block right-hand sides (3 columns each) located at permuted block
// columns cols[k]; returns for each k the full solution gathered at blocks rows[k] (3x3 row-major).
int dev_marginals(DeviceSolver* d, int n, const int* col_p, const int* row_p, double* cov_out,
                  std::string* err);
int dev_label_star(DeviceSolver* d, int gauge_vertex, int n, const int* v, const double* cov,
                   double* meas_out, double* info_out, std::string* err);

}  // namespace pgo
#endif
