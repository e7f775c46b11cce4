/* pgo_solver.h -- C ABI of the B200 SE(2) pose-graph solver (libcgmrslam_b200.so).
 *
 * Drop-in boundary for the arithmetic the reference delegates to g2o ("Seam S", SURVEY.md section
 * 8b): SparseOptimizer::initializeOptimization / optimize(n) with BlockSolver<-1,-1> +
 * LinearSolverCSparse + OptimizationAlgorithmGaussNewton (configured at
 * src/slam/graph_slam.cpp:44-56; called at graph_slam.cpp:392-393,564-565 and
 * graph_manipulator.cpp:116-124), SparseOptimizer::computeMarginals
 * (graph_manipulator.cpp:134-142), computeInitialGuess (graph_manipulator.cpp:122) and
 * EdgeLabeler::labelEdges (condensed_graph_creator.cpp:62-63). A g2o-compatible C++ layer
 * (include/g2o_compat/) marshals the active vertices and edges into the flat arrays below.
 *
 * Conventions
 *   - vertices are the ACTIVE vertices in ascending id order (g2o's activeVertices order,
 *     SURVEY appendix C6), poses are packed (x, y, theta) doubles; edges are EdgeSE2: vertex
 *     indices (i, j) into that array, measurement (dx, dy, dtheta), information as the 6 upper-
 *     triangle entries (I11 I12 I13 I22 I23 I33) -- the layout of the g2o text format;
 *   - the caller owns every host buffer; the handle owns device memory and one CUDA stream;
 *   - return 0 on success, a negative pgo_status on failure; pgo_last_error() has the message;
 *   - there is no CPU fallback: compute entry points need a CUDA device.
 */
#ifndef PGO_SOLVER_H
#define PGO_SOLVER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pgo_status {
  PGO_OK = 0,
  PGO_ERR_ARG = -1,
  PGO_ERR_CUDA = -2,
  PGO_ERR_CAPACITY = -3,
  PGO_ERR_ALLOC = -4,
  PGO_ERR_NUMERIC = -5 /* H not positive definite (g2o: linear solver failure) */
} pgo_status;

typedef struct pgo_solver pgo_solver;

typedef struct pgo_stats {
  int32_t n_vertices, n_edges, n_free;   /* as given / non-fixed vertices                        */
  int32_t n_levels;                      /* elimination-tree height (the kernels run per PANEL level) */
  int64_t factor_blocks;                 /* 3x3 blocks in the factor (incl. diagonal)            */
  int64_t update_ops;                    /* 3x3 block products of one factorisation               */
  int64_t hessian_blocks;                /* non-zero 3x3 blocks of H (lower triangle + diagonal) */
  double analyse_seconds;                /* host time of the structure analysis                  */
  double last_iterate_ms;                /* device time of the last pgo_iterate call             */
  int64_t kernel_launches;               /* kernels launched by this solver so far               */
  double stage_ms[5];                    /* last GN iteration: linearise, factor, forward, backward, update */
  double factor_flops;                   /* fp64 flops of one numeric factorisation (supernodal count) */
  int32_t n_supernodes, n_panels, n_panel_levels, n_supernode_levels;
  int32_t batch;                         /* graph instances held by the solver */
} pgo_stats;

const char* pgo_last_error(void);

/* `stream`: a cudaStream_t to run on, or NULL to create one. */
int pgo_create(pgo_solver** out, int device, void* stream);
void pgo_destroy(pgo_solver* s);

/* SparseOptimizer::initializeOptimization + BlockSolver::buildStructure + the symbolic part of
 * LinearSolverCSparse: active set, hessian indices (fixed -> -1, the others
 * 0..n-1 in vertex order), fill-reducing ordering, factor structure, supernodes / panels, task lists.
 * fixed[v] != 0 marks a fixed vertex. At least one vertex must be fixed per connected component,
 * exactly as with g2o (otherwise H is singular and pgo_iterate reports PGO_ERR_NUMERIC). */
int pgo_set_graph(pgo_solver* s, int n_vertices, int n_edges, const int32_t* edge_i,
                  const int32_t* edge_j, const uint8_t* fixed);

/* Estimates, measurements, information matrices of the graph set above (host -> device). */
int pgo_upload(pgo_solver* s, const double* poses, const double* meas, const double* info6);
/* Estimates only (VertexSE2::setEstimate on every vertex). */
int pgo_set_poses(pgo_solver* s, const double* poses);
int pgo_get_poses(pgo_solver* s, double* poses);

/* SparseOptimizer::optimize(n_iters) with Gauss-Newton: per iteration
 * computeActiveErrors, buildSystem, solve, update; exactly n_iters iterations, no damping, no
 * termination test. chi2_out (may be NULL) receives chi2 at the linearisation point of every
 * iteration; poses_out (may be NULL) the final estimates. *iters_done = iterations completed
 * (fewer than n_iters only if the factorisation fails, in which case the result is
 * PGO_ERR_NUMERIC and the estimates are those before the failing iteration). */
int pgo_iterate(pgo_solver* s, int n_iters, double* poses_out, double* chi2_out, int* iters_done);

/* ---- batch of graphs with one structure --------------------------------------------------------
 * pgo_set_batch(b) (before pgo_set_graph) makes the solver hold b instances of the graph: same
 * vertices, edges and fixed set, separate estimates / measurements / information matrices (e.g.
 * the graphs of several robots sharing a trajectory topology, or resampled measurements). Every
 * kernel of an iteration then serves all instances, which is what fills the GPU: a single
 * 50k-vertex factorisation is latency-bound (about a hundred dependent panel levels). pgo_upload
 * and pgo_set_poses address every instance, the *_instance calls one; pgo_iterate runs all
 * instances and reports instance 0, pgo_iterate_batch reports all: chi2_out[batch][n_iters] (may be
 * NULL), iters_done[batch]. Marginals, edge labelling and the initial guess act on instance 0. */
int pgo_set_batch(pgo_solver* s, int batch);
int pgo_upload_instance(pgo_solver* s, int inst, const double* poses, const double* meas,
                        const double* info6);
int pgo_get_poses_instance(pgo_solver* s, int inst, double* poses);
int pgo_iterate_batch(pgo_solver* s, int n_iters, double* chi2_out, int* iters_done);

/* EdgeSE2::computeError + chi2() over all edges at the current estimates (computeActiveErrors). */
int pgo_chi2(pgo_solver* s, double* chi2);

/* SparseOptimizer::computeMarginals(spinv, blockIndices): blocks (r, c) of H^-1 for
 * the H assembled by the LAST iteration of pgo_iterate (not re-linearised). vertex_r / vertex_c
 * are vertex indices of non-fixed vertices; cov_out[k] is the row-major 3x3 block. */
int pgo_marginals(pgo_solver* s, int n_blocks, const int32_t* vertex_r, const int32_t* vertex_c,
                  double* cov_out);

/* SparseOptimizer::computeInitialGuess: breadth-first propagation of the edge
 * measurements from the fixed vertices (EdgeSE2::initialEstimate); ties broken by hop count, then
 * edge index. Overwrites the estimates of every reached non-fixed vertex. */
int pgo_initial_guess(pgo_solver* s);

/* EdgeLabeler::labelEdges for star edges gauge -> v with the gauge fixed (used by
 * CondensedGraphCreator::compute, condensed_graph_creator.cpp:49-63): measurement = Xg^-1 * Xv
 * at the current estimates, information = inverse of the unscented-transform covariance of the
 * edge error under v's marginal (from the last factorisation). meas_out [n][3], info_out [n][9]. */
int pgo_label_star_edges(pgo_solver* s, int gauge, int n, const int32_t* v, double* meas_out,
                         double* info_out);

/* ---- one graph on several GPUs (domain decomposition) ----------------------
 * One process per GPU, all with the same graph. pgo_set_partition(rank, world) (world a power of
 * two) before pgo_set_graph makes the structure analysis cut the top log2(world) dissection
 * levels into shared separators and give every other vertex to one rank. One iteration is then
 *     pgo_dd_local                       this rank's interior: linearise, eliminate, forward-solve
 *     all-reduce(sum) of the exchange buffer over the ranks    <- the only collective: the
 *                                        separator Schur blocks + right-hand side + chi2
 *     pgo_dd_shared                      separators (replicated), back-substitution, update
 * bracketed by pgo_dd_begin / pgo_dd_end; afterwards pgo_dd_pose_exchange + all-reduce(sum) +
 * pgo_dd_pose_commit leave the complete estimate vector on every rank. The all-reduce is the
 * caller's (ncclAllReduce, or torch.distributed on a tensor aliasing the buffer), enqueued on the
 * solver's stream; everything here is asynchronous on that stream except pgo_dd_end. */
int pgo_set_partition(pgo_solver* s, int rank, int world);
int pgo_dd_begin(pgo_solver* s, int n_iters);
int pgo_dd_local(pgo_solver* s);
int pgo_dd_exchange_buffer(pgo_solver* s, void** dev_ptr, int64_t* n_doubles);
int pgo_dd_shared(pgo_solver* s);
int pgo_dd_end(pgo_solver* s, int n_iters, double* chi2_out, int* iters_done);
int pgo_dd_pose_exchange(pgo_solver* s, void** dev_ptr, int64_t* n_doubles);
int pgo_dd_pose_commit(pgo_solver* s);
/* Host-only (no device needed): the partition the analysis would choose. vertex_owner[v] = rank,
 * -1 = shared separator, -2 = fixed. stats[0] = shared vertices, stats[1] = factor blocks in shared
 * columns, stats[2 + r] = block updates sourced by rank r, stats[2 + world] = by the separators. */
/* The same loop inside the library, with NCCL (pgo_set_partition(handle, vertex_rank,
 * ncclComm_t)): pgo_dd_set_comm adopts the host program's ncclComm_t, pgo_dd_comm_init creates one
 * from a 128-byte ncclUniqueId (pgo_dd_unique_id on rank 0, broadcast by the caller); both imply
 * pgo_set_partition(rank, world) and must precede pgo_set_graph. pgo_dd_iterate = n x (local stage,
 * ncclAllReduce of the separator blocks on the solver's stream, shared stage) + the final exchange
 * of the estimates. libnccl.so.2 is bound at run time (dlopen), not at link time. */
int pgo_dd_unique_id(void* out128);
int pgo_dd_comm_init(pgo_solver* s, const void* unique_id128, int rank, int world);
int pgo_dd_set_comm(pgo_solver* s, void* nccl_comm, int rank, int world);
int pgo_dd_iterate(pgo_solver* s, int n_iters, double* chi2_out, int* iters_done);

int pgo_analyse_partition(int n_vertices, int n_edges, const int32_t* edge_i, const int32_t* edge_j,
                          const uint8_t* fixed, int world, int32_t* vertex_owner, int64_t* stats);

int pgo_get_stats(const pgo_solver* s, pgo_stats* out);
void* pgo_stream(const pgo_solver* s);

#ifdef __cplusplus
}
#endif
#endif
