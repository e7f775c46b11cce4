// TEST INFRASTRUCTURE ONLY (oracle/).
// extern "C" driver over the reference's own CharGrid, compiled verbatim from
// /root/reference/src/matcher/chargrid.cpp by oracle/Makefile into oracle/_ref/libref_chargrid.so.
// It adds no arithmetic of its own: every entry point forwards to a reference member function
// so tests (and bench.py's cpu_baseline / --impl reference legs) can call the real code through
// ctypes. Only tests/, __graft_entry__.smoke() and bench.py may load this library.
#include <cstring>
#include <vector>

#include "chargrid.h"  // the reference header, found via -I/root/reference/src/matcher

namespace {

Vector2dVector to_points(const double* xy, int n) {
  Vector2dVector v(n);
  for (int i = 0; i < n; ++i) v[i] = Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]);
  return v;
}

RegionVector to_regions(const float* r6, int n) {
  RegionVector v(n);
  for (int i = 0; i < n; ++i) {
    v[i].lowerLeft = Eigen::Vector3f(r6[6 * i + 0], r6[6 * i + 1], r6[6 * i + 2]);
    v[i].upperRight = Eigen::Vector3f(r6[6 * i + 3], r6[6 * i + 4], r6[6 * i + 5]);
  }
  return v;
}

int emit(const std::vector<MatcherResult>& res, double* out4, int cap) {
  int n = static_cast<int>(res.size());
  for (int i = 0; i < n && i < cap; ++i) {
    out4[4 * i + 0] = res[i].transformation.x();
    out4[4 * i + 1] = res[i].transformation.y();
    out4[4 * i + 2] = res[i].transformation.z();
    out4[4 * i + 3] = res[i].score;
  }
  return n;
}

}  // namespace

extern "C" {

// CharGrid ctor, chargrid.cpp:124-128 (-> _GridMap ctor gridmap.h:196-214).
void* ref_grid_create(float llx, float lly, float urx, float ury, float res, int kscale) {
  return new CharGrid(Eigen::Vector2f(llx, lly), Eigen::Vector2f(urx, ury), res, kscale);
}

void ref_grid_destroy(void* h) { delete static_cast<CharGrid*>(h); }

void ref_grid_size(void* h, int* rows, int* cols) {
  Eigen::Vector2i s = static_cast<CharGrid*>(h)->grid().size();
  *rows = s.x();
  *cols = s.y();
}

// Same loop as ScanMatcher::resetGrid (scan_matcher.cpp:68-76): every cell <- value.
void ref_grid_fill(void* h, int value) {
  CharGrid* g = static_cast<CharGrid*>(h);
  Eigen::Vector2i s = g->grid().size();
  for (int i = 0; i < s.x(); ++i)
    for (int j = 0; j < s.y(); ++j) g->grid().cell(i, j) = value;
}

// CharGrid::addAndConvolvePoints (chargrid.h:205-216) with a column-major kdim x kdim stamp.
void ref_grid_raster(void* h, const double* xy, int n, const unsigned char* stamp_colmajor,
                     int kdim) {
  CharGrid* g = static_cast<CharGrid*>(h);
  MatrixXChar ker;
  ker.resize(kdim, kdim);
  std::memcpy(ker.data(), stamp_colmajor, static_cast<size_t>(kdim) * kdim);
  Vector2dVector pts = to_points(xy, n);
  g->addAndConvolvePoints<Vector2dVector>(pts.begin(), pts.end(), ker);
}

// Dense copy of the cells, row-major [x][y] (cell(x,y) = rows[x][y], gridmap.h:66-69).
void ref_grid_download(void* h, unsigned char* dst) {
  CharGrid* g = static_cast<CharGrid*>(h);
  Eigen::Vector2i s = g->grid().size();
  for (int i = 0; i < s.x(); ++i)
    for (int j = 0; j < s.y(); ++j) dst[static_cast<size_t>(i) * s.y() + j] = g->grid().cell(i, j);
}

void ref_grid_upload(void* h, const unsigned char* src) {
  CharGrid* g = static_cast<CharGrid*>(h);
  Eigen::Vector2i s = g->grid().size();
  for (int i = 0; i < s.x(); ++i)
    for (int j = 0; j < s.y(); ++j) g->grid().cell(i, j) = src[static_cast<size_t>(i) * s.y() + j];
}

void ref_world2grid(void* h, float x, float y, int* ix, int* iy) {
  Eigen::Vector2i c = static_cast<CharGrid*>(h)->grid().world2grid(Eigen::Vector2f(x, y));
  *ix = c.x();
  *iy = c.y();
}

void ref_grid2world(void* h, int ix, int iy, float* x, float* y) {
  Eigen::Vector2f w = static_cast<CharGrid*>(h)->grid().grid2world(Eigen::Vector2i(ix, iy));
  *x = w.x();
  *y = w.y();
}

// CharGrid::subsample, chargrid.cpp:98-122. dst must hold n points; returns the count.
int ref_subsample(const double* xy, int n, double res, double* dst_xy) {
  Vector2dVector src = to_points(xy, n), dst;
  CharGrid::subsample(dst, src, res);
  for (size_t i = 0; i < dst.size(); ++i) {
    dst_xy[2 * i] = dst[i].x();
    dst_xy[2 * i + 1] = dst[i].y();
  }
  return static_cast<int>(dst.size());
}

// CharGrid::greedySearch(mresvec, points, regions, params), chargrid.cpp:208-308.
// regions: [nreg][6] = lower(x,y,th), upper(x,y,th) as float (Region, chargrid.h:87-92).
// Returns the number of results; the first min(cap, n) are written as (x, y, theta, score).
int ref_greedy_search(void* h, const double* xy, int n, const float* regions, int nreg,
                      double step_x, double step_y, double step_th, double max_score, double dx,
                      double dy, double dth, double* out4, int cap) {
  CharGrid* g = static_cast<CharGrid*>(h);
  Vector2dVector pts = to_points(xy, n);
  RegionVector regs = to_regions(regions, nreg);
  MatchingParameters p;
  p.searchStep = Eigen::Vector3d(step_x, step_y, step_th);
  p.maxScore = max_score;
  p.resultsDiscretization = Eigen::Vector3d(dx, dy, dth);
  std::vector<MatcherResult> res;
  g->greedySearch(res, pts, regs, p);
  return emit(res, out4, cap);
}

// The overload ScanMatcher actually calls for close/LC matching (chargrid.cpp:182-206):
// search step = (grid resolution, grid resolution, thetaRes).
int ref_greedy_search_res(void* h, const double* xy, int n, const float* regions, int nreg,
                          double theta_res, double max_score, double dx, double dy, double dth,
                          double* out4, int cap) {
  CharGrid* g = static_cast<CharGrid*>(h);
  Vector2dVector pts = to_points(xy, n);
  RegionVector regs = to_regions(regions, nreg);
  std::vector<MatcherResult> res;
  g->greedySearch(res, pts, regs, theta_res, max_score, dx, dy, dth);
  return emit(res, out4, cap);
}

// CharGrid::hierarchicalSearch(..., nLevels), chargrid.cpp:376-400 -> :310-344.
int ref_hierarchical_search(void* h, const double* xy, int n, const float* regions, int nreg,
                            double theta_res, double max_score, double dx, double dy, double dth,
                            int n_levels, double* out4, int cap) {
  CharGrid* g = static_cast<CharGrid*>(h);
  Vector2dVector pts = to_points(xy, n);
  RegionVector regs = to_regions(regions, nreg);
  std::vector<MatcherResult> res;
  g->hierarchicalSearch(res, pts, regs, theta_res, max_score, dx, dy, dth, n_levels);
  return emit(res, out4, cap);
}

// CharGrid::countPoints, chargrid.cpp:417-441.
double ref_count_points(void* h, float llx, float lly, float urx, float ury) {
  double score = 0;
  static_cast<CharGrid*>(h)->countPoints(Eigen::Vector2f(llx, lly), Eigen::Vector2f(urx, ury),
                                         &score);
  return score;
}

// CharGrid::searchNonMatchedPoints, chargrid.cpp:444-455. dst must hold n points.
int ref_search_non_matched(void* h, const double* xy, int n, double max_score, double* dst_xy) {
  Vector2dVector pts = to_points(xy, n), out;
  static_cast<CharGrid*>(h)->searchNonMatchedPoints(pts, out, max_score);
  for (size_t i = 0; i < out.size(); ++i) {
    dst_xy[2 * i] = out[i].x();
    dst_xy[2 * i + 1] = out[i].y();
  }
  return static_cast<int>(out.size());
}

}  // extern "C"
