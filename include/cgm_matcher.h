/* cgm_matcher.h -- C ABI of the B200 correlative scan matcher (libcgmrslam_b200.so).
 *
 * Drop-in boundary for the reference's matcher path ("Seam M"). The
 * reference has no FFI: GraphSLAM / MRGraphSLAM call ScanMatcher / CharGrid member functions
 * in-process (src/slam/graph_slam.cpp:244,444,463; src/mrslam/mr_graph_slam.cpp:213,220,287,295).
 * Each entry point below names the reference function it replaces; INTEGRATION.md shows the
 * ScanMatcher / CharGrid shim a maintainer links instead of src/matcher/chargrid.cpp.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer, the handle owns all
 *     device memory and one CUDA stream;
 *   - every function returns 0 on success, a negative cgm_status on failure (the C++ shim maps
 *     <0 to `false` + a line on std::cerr, which is how the reference reports failures);
 *     cgm_last_error() gives the message of the last failure on the calling thread;
 *   - points are packed (x, y) doubles; regions are packed 6-float rows
 *     (lower x, y, theta, upper x, y, theta) exactly like Region (chargrid.h:87-92);
 *   - a matcher owns `n_slots` grids of identical geometry. Slot 0 is "the" CharGrid of the
 *     reference API; the *_batch calls run many (grid, scan, regions) problems in one launch;
 *   - there is no CPU fallback: without a usable CUDA device every compute call fails with
 *     CGM_ERR_CUDA.
 */
#ifndef CGM_MATCHER_H
#define CGM_MATCHER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cgm_status {
  CGM_OK = 0,
  CGM_ERR_ARG = -1,      /* bad argument (null pointer, non-positive step, slot out of range ...) */
  CGM_ERR_CUDA = -2,     /* CUDA runtime error or no device */
  CGM_ERR_CAPACITY = -3, /* a documented size limit was exceeded */
  CGM_ERR_ALLOC = -4
} cgm_status;

typedef struct cgm_matcher cgm_matcher;

/* MatcherResult (chargrid.h:50-60) without the unused information matrix. */
typedef struct cgm_result {
  double x, y, theta, score;
} cgm_result;

const char* cgm_last_error(void);
int cgm_device_count(void);
/* Number of kernels this library has launched in the calling process (bench.py "gpu_launches"). */
uint64_t cgm_launch_count(void);

/* CharGrid ctor (chargrid.cpp:124-128 -> _GridMap ctor gridmap.h:196-214) + ScanMatcher::
 * initializeKernel (scan_matcher.cpp:38-61) + initializeGrid (:63-66), for n_slots grids on
 * `device`. `stream` is a cudaStream_t to run on, or NULL to let the matcher create its own.
 * `resolution` is the double both reference calls receive: the stamp is built from it as a
 * double (scan_matcher.cpp:39,45), the grid narrows it to float (scan_matcher.cpp:64). */
int cgm_matcher_create(cgm_matcher** out, int device, void* stream, int n_slots, float llx,
                       float lly, float urx, float ury, double resolution, double kernel_range,
                       int kscale);
void cgm_matcher_destroy(cgm_matcher* m);

/* _GridMap::size() (gridmap.h:134): rows = x extent, cols = y extent. */
int cgm_matcher_grid_size(const cgm_matcher* m, int* rows, int* cols);
/* The distance stamp built by initializeKernel, column-major dim x dim. */
int cgm_matcher_stamp(const cgm_matcher* m, uint8_t* dst, int cap, int* dim);
/* _GridMap::world2grid (gridmap.h:24-27) / grid2world (:45-48): host-side float arithmetic. */
int cgm_matcher_world2grid(const cgm_matcher* m, float x, float y, int* ix, int* iy);
int cgm_matcher_grid2world(const cgm_matcher* m, int ix, int iy, float* x, float* y);

/* ScanMatcher::resetGrid (scan_matcher.cpp:68-76) on one slot: every cell <- int(range*kscale). */
int cgm_matcher_reset(cgm_matcher* m, int slot);
/* CharGrid::addAndConvolvePoints (chargrid.h:205-216 -> applyKernel chargrid.cpp:132-161). */
int cgm_matcher_raster(cgm_matcher* m, int slot, const double* map_xy, int n);
/* The kernel argument of CharGrid::addAndConvolvePoints / applyKernel (chargrid.h:190,205-216): the
 * reference's callers build the stamp themselves (ScanMatcher::initializeKernel,
 * scan_matcher.cpp:38-61) and pass it with every call; the CharGrid shim forwards it here when it
 * changes. Column-major dim x dim bytes, dim odd. */
int cgm_matcher_set_stamp(cgm_matcher* m, const uint8_t* stamp_colmajor, int dim);
/* Every cell <- value: the loop of ScanMatcher::resetGrid (scan_matcher.cpp:68-76) for any fill
 * value; with points, followed by addAndConvolvePoints in the same launch. */
int cgm_matcher_fill(cgm_matcher* m, int slot, int value);
int cgm_matcher_fill_raster(cgm_matcher* m, int slot, int value, const double* map_xy, int n);
/* CharGrid's copy constructor / assignment (`CharGrid auxGrid = _grid;`, scan_matcher.cpp:437):
 * device-to-device copy of one slot's cells between matchers of the same geometry and device. */
int cgm_matcher_copy_grid(cgm_matcher* dst, int dst_slot, cgm_matcher* src, int src_slot);
/* Dense copy of the cells, row-major [x][y] (cell(x,y) = rows[x][y], gridmap.h:66-69);
 * serves ScanMatcher::grid() (scan_matcher.h:77) and the parity tests. */
int cgm_matcher_grid_download(cgm_matcher* m, int slot, uint8_t* dst);
int cgm_matcher_grid_upload(cgm_matcher* m, int slot, const uint8_t* src);

/* CharGrid::subsample (chargrid.cpp:98-122). Host-side (ordered-map semantics, O(n log n) on a
 * few hundred points); dst must hold n points. */
int cgm_subsample(const double* src_xy, int n, double res, double* dst_xy, int* n_out);

/* CharGrid::greedySearch(mresvec, points, regions, params) (chargrid.cpp:208-308) on one slot.
 * Writes min(*n_out, cap) results sorted like the reference's mresvec; *n_out is the full count. */
int cgm_matcher_search(cgm_matcher* m, int slot, const double* pts_xy, int n_pts,
                       const float* regions, int n_regions, double step_x, double step_y,
                       double step_theta, double max_score, double bin_x, double bin_y,
                       double bin_theta, cgm_result* out, int cap, int* n_out);

/* CharGrid::hierarchicalSearch(mresvec, points, regions, thetaRes, maxScore, dx, dy, dth, nLevels)
 * (chargrid.cpp:376-400 -> :310-344). */
int cgm_matcher_hierarchical_search(cgm_matcher* m, int slot, const double* pts_xy, int n_pts,
                                    const float* regions, int n_regions, double theta_res,
                                    double max_score, double bin_x, double bin_y, double bin_theta,
                                    int n_levels, cgm_result* out, int cap, int* n_out);

/* CharGrid::hierarchicalSearch(mresvec, points, regions, paramsVec) (chargrid.cpp:310-344) with the
 * caller's own ladder: levels7 = n_levels rows of (step x, y, theta, maxScore, bin x, y, theta). */
int cgm_matcher_hierarchical_search_levels(cgm_matcher* m, int slot, const double* pts_xy, int n_pts,
                                           const float* regions, int n_regions, const double* levels7,
                                           int n_levels, cgm_result* out, int cap, int* n_out);

/* CharGrid::countPoints (chargrid.cpp:417-441). */
int cgm_matcher_count_points(cgm_matcher* m, int slot, float llx, float lly, float urx, float ury,
                             double* score);
/* CharGrid::searchNonMatchedPoints (chargrid.cpp:444-455); dst must hold n points. */
int cgm_matcher_search_non_matched(cgm_matcher* m, int slot, const double* pts_xy, int n,
                                   double max_score, double* dst_xy, int* n_out);

/* ---- batch: slots [first_slot, first_slot + n) processed together ------------------------ */

/* resetGrid + addAndConvolvePoints for n slots. map_xy packs the slots' points back to back;
 * counts[s] is the number of points of slot first_slot + s. */
int cgm_matcher_raster_batch(cgm_matcher* m, int first_slot, int n, const double* map_xy,
                             const int* counts);

/* n independent greedySearch problems (slot s: its grid, its points, its regions), common search
 * parameters, one scoring launch. out holds n * cap_per_slot results; n_out[s] is slot s's full
 * result count. This is also the e2e path of bench.py: host buffers in, host results out. */
int cgm_matcher_search_batch(cgm_matcher* m, int first_slot, int n, const double* pts_xy,
                             const int* pts_counts, const float* regions, const int* region_counts,
                             double step_x, double step_y, double step_theta, double max_score,
                             double bin_x, double bin_y, double bin_theta, cgm_result* out,
                             int cap_per_slot, int* n_out);

/* The same work split in three so that a harness can time the device part alone with inputs
 * already resident in HBM (bench.py "value"):
 *   stage   -- host planning + H2D of points, regions, theta tables (and map points if given);
 *   launch  -- kernels only, asynchronous on the matcher's stream: [reset + raster if map points
 *              were staged] + rotate/quantise + score + compact;
 *   collect -- stream sync + D2H of survivors + host-side bin merge and std::sort. */
int cgm_matcher_batch_stage(cgm_matcher* m, int first_slot, int n, const double* map_xy,
                            const int* map_counts, const double* pts_xy, const int* pts_counts,
                            const float* regions, const int* region_counts, double step_x,
                            double step_y, double step_theta, double max_score, double bin_x,
                            double bin_y, double bin_theta);
int cgm_matcher_batch_launch(cgm_matcher* m);
int cgm_matcher_batch_collect(cgm_matcher* m, cgm_result* out, int cap_per_slot, int* n_out);

/* Bookkeeping of the last staged batch, for roofline accounting:
 * candidates = number of scores, cell_reads = sum over candidates of k_theta (known after
 * collect), score_launches = scoring kernel launches of the last launch call. */
int cgm_matcher_batch_stats(const cgm_matcher* m, uint64_t* candidates, uint64_t* cell_reads,
                            int* score_launches);
/* Device time in ms of the last batch launch, measured with CUDA events on the matcher's stream:
 * ms[0] reset+raster kernels, ms[1] the scoring kernel, ms[2] compaction. Synchronises. */
int cgm_matcher_batch_kernel_ms(cgm_matcher* m, float* ms3);
/* The cudaStream_t the matcher launches on (so a harness can record events on it). */
void* cgm_matcher_stream(const cgm_matcher* m);
/* Select the scoring kernel: 0 = auto, 1 = global-memory reference kernel, 2 = shared-memory
 * tiled kernel (fails with CGM_ERR_CAPACITY at search time if a tile does not fit). */
int cgm_matcher_set_kernel(cgm_matcher* m, int which);

#ifdef __cplusplus
}
#endif
#endif
