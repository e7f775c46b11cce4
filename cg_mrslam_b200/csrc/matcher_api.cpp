// extern "C" surface of the scan matcher (include/cgm_matcher.h): argument checking, host-side
// planning (matcher_plan.cpp), device work (matcher_device.h) and result assembly.
#include <algorithm>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/cgm_matcher.h"
#include "matcher_device.h"
#include "matcher_plan.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

}  // namespace

struct cgm_matcher {
  cgm::GridGeom geom;
  int n_slots;
  std::vector<uint8_t> stamp;
  int stamp_dim;
  double kernel_range;
  cgm::DeviceMatcher* dev;
  int kernel_choice;
  // last staged batch
  cgm::SearchPlan plan;
  bool staged, map_staged, launched;
  uint64_t cell_reads;
  int score_launches;
};

namespace {

bool slot_ok(const cgm_matcher* m, int first, int n) {
  return first >= 0 && n >= 0 && first + n <= m->n_slots;
}

void write_results(const std::vector<cgm_result>& res, cgm_result* out, int cap, int* n_out) {
  const int n = static_cast<int>(res.size());
  for (int i = 0; i < n && i < cap; ++i) out[i] = res[i];
  if (n_out) *n_out = n;
}

// One greedySearch over n problems, host buffers in, per-problem result vectors out.
int run_search(cgm_matcher* m, int first_slot, int n, const double* pts_xy, const int* pts_counts,
               const float* regions, const int* region_counts, const cgm::SearchParams& p,
               std::vector<std::vector<cgm_result> >* results) {
  std::string err;
  cgm::SearchPlan plan;
  int rc = cgm::build_plan(m->geom, first_slot, n, pts_xy, pts_counts, regions, region_counts, p,
                           &plan, &err);
  if (rc) return fail(rc, err);
  int total_pts = 0;
  for (int i = 0; i < n; ++i) total_pts += pts_counts[i];
  std::vector<cgm::Survivor> survivors;
  if (!plan.units.empty()) {
    rc = cgm::dev_stage_search(m->dev, plan, pts_xy, total_pts, &err);
    if (rc) return fail(rc, err);
    int launches = 0;
    rc = cgm::dev_launch_search(m->dev, plan, m->kernel_choice, &launches, &err);
    if (rc) return fail(rc, err);
    uint64_t reads = 0;
    rc = cgm::dev_collect(m->dev, plan, &survivors, &reads, &err);
    if (rc) return fail(rc, err);
  }
  cgm::assemble_results(m->geom, plan, survivors, results);
  return CGM_OK;
}

}  // namespace

extern "C" {

const char* cgm_last_error(void) { return g_last_error.c_str(); }

int cgm_device_count(void) { return cgm::dev_device_count(); }

uint64_t cgm_launch_count(void) { return cgm::dev_launch_count(); }

int cgm_matcher_create(cgm_matcher** out, int device, void* stream, int n_slots, float llx,
                       float lly, float urx, float ury, double resolution, double kernel_range,
                       int kscale) {
  if (!out) return fail(CGM_ERR_ARG, "out is null");
  *out = nullptr;
  if (n_slots <= 0) return fail(CGM_ERR_ARG, "n_slots must be positive");
  if (!(resolution > 0.0) || !(static_cast<float>(resolution) > 0.f)) return fail(CGM_ERR_ARG, "resolution must be positive");
  if (!(kernel_range >= 0.0) || kscale <= 0) return fail(CGM_ERR_ARG, "bad kernel range / scale");
  cgm_matcher* m = new (std::nothrow) cgm_matcher();
  if (!m) return fail(CGM_ERR_ALLOC, "out of host memory");
  m->geom = cgm::make_geom(llx, lly, urx, ury, static_cast<float>(resolution), kernel_range, kscale);
  if (m->geom.rows <= 0 || m->geom.cols <= 0) {
    delete m;
    return fail(CGM_ERR_ARG, "empty grid");
  }
  m->n_slots = n_slots;
  m->kernel_range = kernel_range;
  m->stamp = cgm::make_stamp(resolution, kernel_range, kscale, &m->stamp_dim);
  if (m->stamp_dim > 63) {  // the rasteriser tabulates the stamp's cells in shared memory
    delete m;
    return fail(CGM_ERR_CAPACITY, "distance stamp wider than 63 cells (kernel_range / resolution > 31)");
  }
  // Cells only ever hold the fill value or a stamp value; the kernels size their packed
  // accumulators from this bound (a raw grid upload raises it to 255).
  for (size_t i = 0; i < m->stamp.size(); ++i)
    m->geom.max_cell = std::max<int>(m->geom.max_cell, m->stamp[i]);
  m->kernel_choice = 0;
  m->staged = m->map_staged = m->launched = false;
  m->cell_reads = 0;
  m->score_launches = 0;
  std::string err;
  int rc = cgm::dev_create(&m->dev, device, stream, n_slots, m->geom, m->stamp.data(),
                           m->stamp_dim, &err);
  if (rc) {
    delete m;
    return fail(rc, err);
  }
  *out = m;
  return CGM_OK;
}

void cgm_matcher_destroy(cgm_matcher* m) {
  if (!m) return;
  cgm::dev_destroy(m->dev);
  delete m;
}

int cgm_matcher_grid_size(const cgm_matcher* m, int* rows, int* cols) {
  if (!m || !rows || !cols) return fail(CGM_ERR_ARG, "null argument");
  *rows = m->geom.rows;
  *cols = m->geom.cols;
  return CGM_OK;
}

int cgm_matcher_stamp(const cgm_matcher* m, uint8_t* dst, int cap, int* dim) {
  if (!m || !dim) return fail(CGM_ERR_ARG, "null argument");
  *dim = m->stamp_dim;
  if (dst) {
    if (cap < m->stamp_dim * m->stamp_dim) return fail(CGM_ERR_ARG, "stamp buffer too small");
    std::memcpy(dst, m->stamp.data(), m->stamp.size());
  }
  return CGM_OK;
}

int cgm_matcher_world2grid(const cgm_matcher* m, float x, float y, int* ix, int* iy) {
  if (!m || !ix || !iy) return fail(CGM_ERR_ARG, "null argument");
  cgm::world2grid(m->geom, x, y, ix, iy);
  return CGM_OK;
}

int cgm_matcher_grid2world(const cgm_matcher* m, int ix, int iy, float* x, float* y) {
  if (!m || !x || !y) return fail(CGM_ERR_ARG, "null argument");
  cgm::grid2world(m->geom, ix, iy, x, y);
  return CGM_OK;
}

int cgm_matcher_reset(cgm_matcher* m, int slot) {
  if (!m || !slot_ok(m, slot, 1)) return fail(CGM_ERR_ARG, "bad slot");
  std::string err;
  int zero = 0;
  int rc = cgm::dev_stage_map(m->dev, slot, 1, nullptr, &zero, true, &err);
  if (!rc) rc = cgm::dev_launch_map(m->dev, &err);
  if (!rc) rc = cgm::dev_sync(m->dev, &err);
  cgm::dev_clear_map_stage(m->dev);
  return rc ? fail(rc, err) : CGM_OK;
}

int cgm_matcher_raster(cgm_matcher* m, int slot, const double* map_xy, int n) {
  if (!m || !slot_ok(m, slot, 1) || n < 0 || (n && !map_xy)) return fail(CGM_ERR_ARG, "bad argument");
  std::string err;
  int rc = cgm::dev_stage_map(m->dev, slot, 1, map_xy, &n, false, &err);
  if (!rc) rc = cgm::dev_launch_map(m->dev, &err);
  if (!rc) rc = cgm::dev_sync(m->dev, &err);
  cgm::dev_clear_map_stage(m->dev);
  return rc ? fail(rc, err) : CGM_OK;
}

int cgm_matcher_set_stamp(cgm_matcher* m, const uint8_t* stamp_colmajor, int dim) {
  if (!m || !stamp_colmajor || dim <= 0 || dim % 2 == 0 || dim > 63)
    return fail(CGM_ERR_ARG, "bad stamp (odd dimension 1..63 expected)");
  std::string err;
  const int rc = cgm::dev_set_stamp(m->dev, stamp_colmajor, dim, &err);
  if (rc) return fail(rc, err);
  m->stamp.assign(stamp_colmajor, stamp_colmajor + static_cast<size_t>(dim) * dim);
  m->stamp_dim = dim;
  // stamping takes minima: a cell never exceeds what it held before, so the bound only has to
  // cover cells written by fills / uploads; it is kept as it is
  return CGM_OK;
}

namespace {
// every cell of `slot` <- value, then the points are stamped: one launch
int fill_and_raster(cgm_matcher* m, int slot, int value, const double* map_xy, int n) {
  std::string err;
  // cells hold `value` or less from now on; other slots of a multi-slot matcher keep their bound
  const int bound = m->n_slots == 1 ? value : std::max(m->geom.max_cell, value);
  m->geom.fill_value = value;
  m->geom.max_cell = std::max(bound, 1);
  cgm::dev_set_bounds(m->dev, value, m->geom.max_cell);
  int rc = cgm::dev_stage_map(m->dev, slot, 1, map_xy, &n, true, &err);
  if (!rc) rc = cgm::dev_launch_map(m->dev, &err);
  if (!rc) rc = cgm::dev_sync(m->dev, &err);
  cgm::dev_clear_map_stage(m->dev);
  return rc ? fail(rc, err) : CGM_OK;
}
}  // namespace

int cgm_matcher_fill(cgm_matcher* m, int slot, int value) {
  if (!m || !slot_ok(m, slot, 1) || value < 0 || value > 255) return fail(CGM_ERR_ARG, "bad argument");
  return fill_and_raster(m, slot, value, nullptr, 0);
}

int cgm_matcher_fill_raster(cgm_matcher* m, int slot, int value, const double* map_xy, int n) {
  if (!m || !slot_ok(m, slot, 1) || value < 0 || value > 255 || n < 0 || (n && !map_xy))
    return fail(CGM_ERR_ARG, "bad argument");
  return fill_and_raster(m, slot, value, map_xy, n);
}

int cgm_matcher_copy_grid(cgm_matcher* dst, int dst_slot, cgm_matcher* src, int src_slot) {
  if (!dst || !src || !slot_ok(dst, dst_slot, 1) || !slot_ok(src, src_slot, 1))
    return fail(CGM_ERR_ARG, "bad argument");
  if (dst->geom.rows != src->geom.rows || dst->geom.cols != src->geom.cols)
    return fail(CGM_ERR_ARG, "grid copy between different geometries");
  std::string err;
  const int rc = cgm::dev_copy_grid(dst->dev, dst_slot, src->dev, src_slot, &err);
  if (rc) return fail(rc, err);
  dst->geom.max_cell = std::max(dst->geom.max_cell, src->geom.max_cell);
  cgm::dev_set_bounds(dst->dev, dst->geom.fill_value, dst->geom.max_cell);
  return CGM_OK;
}

int cgm_matcher_raster_batch(cgm_matcher* m, int first_slot, int n, const double* map_xy,
                             const int* counts) {
  if (!m || !slot_ok(m, first_slot, n) || !counts) return fail(CGM_ERR_ARG, "bad argument");
  std::string err;
  int rc = cgm::dev_stage_map(m->dev, first_slot, n, map_xy, counts, true, &err);
  if (!rc) rc = cgm::dev_launch_map(m->dev, &err);
  if (!rc) rc = cgm::dev_sync(m->dev, &err);
  cgm::dev_clear_map_stage(m->dev);
  return rc ? fail(rc, err) : CGM_OK;
}

int cgm_matcher_grid_download(cgm_matcher* m, int slot, uint8_t* dst) {
  if (!m || !slot_ok(m, slot, 1) || !dst) return fail(CGM_ERR_ARG, "bad argument");
  std::string err;
  int rc = cgm::dev_grid_download(m->dev, slot, dst, &err);
  return rc ? fail(rc, err) : CGM_OK;
}

int cgm_matcher_grid_upload(cgm_matcher* m, int slot, const uint8_t* src) {
  if (!m || !slot_ok(m, slot, 1) || !src) return fail(CGM_ERR_ARG, "bad argument");
  std::string err;
  int rc = cgm::dev_grid_upload(m->dev, slot, src, &err);
  if (rc) return fail(rc, err);
  m->geom.max_cell = 255;  // arbitrary bytes may now be present
  return CGM_OK;
}

int cgm_subsample(const double* src_xy, int n, double res, double* dst_xy, int* n_out) {
  if (n < 0 || (n && (!src_xy || !dst_xy)) || !n_out) return fail(CGM_ERR_ARG, "bad argument");
  if (!(res > 0.0)) return fail(CGM_ERR_ARG, "resolution must be positive");
  *n_out = cgm::subsample(src_xy, n, res, dst_xy);
  return CGM_OK;
}

int cgm_matcher_search(cgm_matcher* m, int slot, const double* pts_xy, int n_pts,
                       const float* regions, int n_regions, double step_x, double step_y,
                       double step_theta, double max_score, double bin_x, double bin_y,
                       double bin_theta, cgm_result* out, int cap, int* n_out) {
  if (!m || !slot_ok(m, slot, 1) || n_pts < 0 || n_regions < 0 || (n_pts && !pts_xy) ||
      (n_regions && !regions) || cap < 0 || (cap && !out))
    return fail(CGM_ERR_ARG, "bad argument");
  cgm::SearchParams p = {step_x, step_y, step_theta, max_score, bin_x, bin_y, bin_theta};
  std::vector<std::vector<cgm_result> > res;
  int rc = run_search(m, slot, 1, pts_xy, &n_pts, regions, &n_regions, p, &res);
  if (rc) return rc;
  write_results(res[0], out, cap, n_out);
  return CGM_OK;
}

namespace {
// chargrid.cpp:310-344: every level but the last searches the current regions and turns its
// results into the next level's regions; the last level's results are returned.
int run_hierarchy(cgm_matcher* m, int slot, const double* pts_xy, int n_pts, const float* regions,
                  int n_regions, const std::vector<cgm::SearchParams>& levels, cgm_result* out, int cap,
                  int* n_out) {
  std::vector<float> cur(regions, regions + 6 * static_cast<size_t>(n_regions));
  std::vector<cgm_result> last;
  for (size_t li = 0; li + 1 < levels.size(); ++li) {
    int nr = static_cast<int>(cur.size() / 6);
    std::vector<std::vector<cgm_result> > res;
    int rc = run_search(m, slot, 1, pts_xy, &n_pts, cur.data(), &nr, levels[li], &res);
    if (rc) return rc;
    last.swap(res[0]);
    if (last.empty()) break;
    cur = cgm::refine_regions(last, levels[li]);
  }
  // Final pass only if the previous level found something (chargrid.cpp:336). With a single
  // level the loop above does not run, mresvec is still empty and the reference returns nothing.
  if (!last.empty()) {
    int nr = static_cast<int>(cur.size() / 6);
    std::vector<std::vector<cgm_result> > res;
    int rc = run_search(m, slot, 1, pts_xy, &n_pts, cur.data(), &nr, levels.back(), &res);
    if (rc) return rc;
    last.swap(res[0]);
  }
  write_results(last, out, cap, n_out);
  return CGM_OK;
}
}  // namespace

int cgm_matcher_hierarchical_search(cgm_matcher* m, int slot, const double* pts_xy, int n_pts,
                                    const float* regions, int n_regions, double theta_res,
                                    double max_score, double bin_x, double bin_y, double bin_theta,
                                    int n_levels, cgm_result* out, int cap, int* n_out) {
  if (!m || !slot_ok(m, slot, 1) || n_pts < 0 || n_regions < 0 || (n_pts && !pts_xy) ||
      (n_regions && !regions) || cap < 0 || (cap && !out) || n_levels < 1)
    return fail(CGM_ERR_ARG, "bad argument");
  // chargrid.cpp:376-400
  return run_hierarchy(m, slot, pts_xy, n_pts, regions, n_regions,
                       cgm::hierarchical_levels(m->geom, theta_res, max_score, bin_x, bin_y, bin_theta, n_levels),
                       out, cap, n_out);
}

int cgm_matcher_hierarchical_search_levels(cgm_matcher* m, int slot, const double* pts_xy, int n_pts,
                                           const float* regions, int n_regions, const double* levels7,
                                           int n_levels, cgm_result* out, int cap, int* n_out) {
  if (!m || !slot_ok(m, slot, 1) || n_pts < 0 || n_regions < 0 || (n_pts && !pts_xy) ||
      (n_regions && !regions) || cap < 0 || (cap && !out) || n_levels < 1 || !levels7)
    return fail(CGM_ERR_ARG, "bad argument");
  std::vector<cgm::SearchParams> levels(n_levels);
  for (int l = 0; l < n_levels; ++l) {
    const double* p = levels7 + 7 * l;
    const cgm::SearchParams sp = {p[0], p[1], p[2], p[3], p[4], p[5], p[6]};
    levels[l] = sp;
  }
  return run_hierarchy(m, slot, pts_xy, n_pts, regions, n_regions, levels, out, cap, n_out);
}

int cgm_matcher_count_points(cgm_matcher* m, int slot, float llx, float lly, float urx, float ury,
                             double* score) {
  if (!m || !slot_ok(m, slot, 1) || !score) return fail(CGM_ERR_ARG, "bad argument");
  int ax, ay, bx, by;
  cgm::world2grid(m->geom, llx, lly, &ax, &ay);  // chargrid.cpp:422-423
  cgm::world2grid(m->geom, urx, ury, &bx, &by);
  long long sum = 0;
  std::string err;
  int rc = cgm::dev_window_sum(m->dev, slot, ax, ay, bx, by, &sum, &err);
  if (rc) return fail(rc, err);
  const int visited = (bx - ax) * (by - ay);           // chargrid.cpp:434
  const int isum = static_cast<int>(sum);
  *score = static_cast<float>(isum) / static_cast<float>(visited);  // :439
  return CGM_OK;
}

int cgm_matcher_search_non_matched(cgm_matcher* m, int slot, const double* pts_xy, int n,
                                   double max_score, double* dst_xy, int* n_out) {
  if (!m || !slot_ok(m, slot, 1) || n < 0 || (n && (!pts_xy || !dst_xy)) || !n_out)
    return fail(CGM_ERR_ARG, "bad argument");
  std::vector<int> gxy(2 * static_cast<size_t>(n)), val(n);
  for (int i = 0; i < n; ++i)  // chargrid.cpp:448
    cgm::world2grid(m->geom, static_cast<float>(pts_xy[2 * i]), static_cast<float>(pts_xy[2 * i + 1]),
                    &gxy[2 * i], &gxy[2 * i + 1]);
  std::string err;
  if (n) {
    int rc = cgm::dev_cells_at(m->dev, slot, gxy.data(), n, val.data(), &err);
    if (rc) return fail(rc, err);
  }
  const float ikscale = static_cast<float>(1. / static_cast<float>(m->geom.kscale));  // :446
  int k = 0;
  for (int i = 0; i < n; ++i) {
    if (val[i] < 0) continue;  // outside the grid (:449)
    const double value = static_cast<float>(static_cast<uint8_t>(val[i])) * ikscale;  // :450
    if (value > max_score) {
      dst_xy[2 * k] = pts_xy[2 * i];
      dst_xy[2 * k + 1] = pts_xy[2 * i + 1];
      ++k;
    }
  }
  *n_out = k;
  return CGM_OK;
}

int cgm_matcher_search_batch(cgm_matcher* m, int first_slot, int n, const double* pts_xy,
                             const int* pts_counts, const float* regions, const int* region_counts,
                             double step_x, double step_y, double step_theta, double max_score,
                             double bin_x, double bin_y, double bin_theta, cgm_result* out,
                             int cap_per_slot, int* n_out) {
  int rc = cgm_matcher_batch_stage(m, first_slot, n, nullptr, nullptr, pts_xy, pts_counts, regions,
                                   region_counts, step_x, step_y, step_theta, max_score, bin_x,
                                   bin_y, bin_theta);
  if (!rc) rc = cgm_matcher_batch_launch(m);
  if (!rc) rc = cgm_matcher_batch_collect(m, out, cap_per_slot, n_out);
  return rc;
}

int cgm_matcher_batch_stage(cgm_matcher* m, int first_slot, int n, const double* map_xy,
                            const int* map_counts, const double* pts_xy, const int* pts_counts,
                            const float* regions, const int* region_counts, double step_x,
                            double step_y, double step_theta, double max_score, double bin_x,
                            double bin_y, double bin_theta) {
  if (!m || !slot_ok(m, first_slot, n) || !pts_counts || !region_counts)
    return fail(CGM_ERR_ARG, "bad argument");
  m->staged = m->launched = m->map_staged = false;
  cgm::dev_clear_map_stage(m->dev);
  cgm::SearchParams p = {step_x, step_y, step_theta, max_score, bin_x, bin_y, bin_theta};
  std::string err;
  int rc = cgm::build_plan(m->geom, first_slot, n, pts_xy, pts_counts, regions, region_counts, p,
                           &m->plan, &err);
  if (rc) return fail(rc, err);
  if (map_counts) {
    rc = cgm::dev_stage_map(m->dev, first_slot, n, map_xy, map_counts, true, &err);
    if (rc) return fail(rc, err);
    m->map_staged = true;
  }
  int total_pts = 0;
  for (int i = 0; i < n; ++i) total_pts += pts_counts[i];
  if (!m->plan.units.empty()) {
    rc = cgm::dev_stage_search(m->dev, m->plan, pts_xy, total_pts, &err);
    if (rc) return fail(rc, err);
  }
  m->staged = true;
  return CGM_OK;
}

int cgm_matcher_batch_launch(cgm_matcher* m) {
  if (!m || !m->staged) return fail(CGM_ERR_ARG, "nothing staged");
  std::string err;
  int rc = CGM_OK;
  if (m->map_staged) rc = cgm::dev_launch_map(m->dev, &err);
  m->score_launches = 0;
  if (!rc && !m->plan.units.empty())
    rc = cgm::dev_launch_search(m->dev, m->plan, m->kernel_choice, &m->score_launches, &err);
  if (rc) return fail(rc, err);
  m->launched = true;
  return CGM_OK;
}

int cgm_matcher_batch_collect(cgm_matcher* m, cgm_result* out, int cap_per_slot, int* n_out) {
  if (!m || !m->launched) return fail(CGM_ERR_ARG, "nothing launched");
  if (cap_per_slot < 0 || (cap_per_slot && !out)) return fail(CGM_ERR_ARG, "bad output buffer");
  std::string err;
  std::vector<cgm::Survivor> survivors;
  m->cell_reads = 0;
  int rc;
  if (!m->plan.units.empty())
    rc = cgm::dev_collect(m->dev, m->plan, &survivors, &m->cell_reads, &err);
  else
    rc = cgm::dev_sync(m->dev, &err);
  if (rc) return fail(rc, err);
  std::vector<std::vector<cgm_result> > res;
  cgm::assemble_results(m->geom, m->plan, survivors, &res);
  for (int s = 0; s < m->plan.n_problems; ++s)
    write_results(res[s], out ? out + static_cast<size_t>(s) * cap_per_slot : nullptr,
                  cap_per_slot, n_out ? n_out + s : nullptr);
  m->launched = false;  // the bins were drained by the compaction; launch again to re-run
  return CGM_OK;
}

int cgm_matcher_batch_stats(const cgm_matcher* m, uint64_t* candidates, uint64_t* cell_reads,
                            int* score_launches) {
  if (!m) return fail(CGM_ERR_ARG, "null matcher");
  if (candidates) *candidates = m->plan.candidates;
  if (cell_reads) *cell_reads = m->cell_reads;
  if (score_launches) *score_launches = m->score_launches;
  return CGM_OK;
}

int cgm_matcher_batch_kernel_ms(cgm_matcher* m, float* ms3) {
  if (!m || !ms3) return fail(CGM_ERR_ARG, "null argument");
  std::string err;
  int rc = cgm::dev_last_timings(m->dev, ms3, &err);
  return rc ? fail(rc, err) : CGM_OK;
}

void* cgm_matcher_stream(const cgm_matcher* m) { return m ? cgm::dev_stream(m->dev) : nullptr; }

int cgm_matcher_set_kernel(cgm_matcher* m, int which) {
  if (!m || which < 0 || which > 2) return fail(CGM_ERR_ARG, "kernel selector must be 0, 1 or 2");
  m->kernel_choice = which;
  return CGM_OK;
}

}  // extern "C"
