// Host-side planning and result assembly of the scan matcher (no CUDA in this header).
//
// The reference runs CharGrid::greedySearch (src/matcher/chargrid.cpp:208-308) as nested loops
// regions -> theta -> i -> j -> points on the CPU. Here the host only does what must come from
// the host's libm / float unit to stay bit-identical (theta ladder, cos/sin, world<->grid float
// maths, result-bin coordinates) and what is inherently sequential (per-chunk ordered maps,
// std::sort); the O(candidates x points) work is described by flat descriptor tables and done by
// the kernels in matcher_kernels.cu.
#ifndef CGM_MATCHER_PLAN_H
#define CGM_MATCHER_PLAN_H

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/cgm_matcher.h"

namespace cgm {

// Geometry of the CharGridMap (gridmap.h:180-189) plus the device layout.
struct GridGeom {
  float llx, lly, urx, ury;
  float res, inv_res;
  int rows, cols;  // rows = x extent, cols = y extent; cell(x, y) at [x * pitch + y]
  int pitch;       // bytes per row on the device: cols rounded up to 16, padding cells are 0
  int kscale;
  int fill_value;  // int(kernel_range * kscale) as written by resetGrid (scan_matcher.cpp:69)
  int max_cell;    // upper bound of any cell value (fill_value, or 255 after a raw upload)
};

// One search region of one problem, as the scoring kernels see it. 64 bytes.
struct RegionDesc {
  int slot;              // which grid
  int pts_off, pts_n;    // the problem's points inside the packed point array
  int llx, lly;          // first candidate offset in cells (world2grid of the region's lower corner)
  int nx, ny;            // candidates per axis
  int xs, ys;            // candidate stride in cells
  int binx_off, biny_off;  // into the bin-coordinate table (relative bin index per i / j step)
  int nby, nbth;         // extents of this region's result-bin box (y, theta); x extent implied
  uint32_t bin_base;     // first entry of the box in the bin array
  int theta_off;         // index of this region's first unit
  int pad;
};

// One (region, theta) work unit. 24 bytes.
struct ThetaDesc {
  double c, s;  // cos/sin from the host's libm (chargrid.cpp:241)
  int region;
  int bin_th;   // theta bin of this unit relative to the region's box
};

// A surviving result bin as the compaction kernel reports it.
struct Survivor {
  uint32_t entry;  // index into the bin array
  uint32_t pad;
  uint64_t key;    // (float bits of score << 32) | candidate index inside the region
};

static const uint64_t kEmptyBin = ~0ull;

struct SearchParams {
  double step_x, step_y, step_theta;
  double max_score;
  double bin_x, bin_y, bin_theta;
};

// Host-only facts about a region needed to decode survivors.
struct RegionHost {
  int problem;
  int chunk;        // OpenMP chunk of the reference (chargrid.cpp:223-232)
  int n_theta;
  uint32_t n_bins;
};

struct SearchPlan {
  SearchParams params;
  int n_problems = 0;
  std::vector<int> problem_slot;
  std::vector<int> problem_first_region, problem_n_regions, problem_n_chunks;
  std::vector<RegionDesc> regions;
  std::vector<RegionHost> regions_host;
  std::vector<ThetaDesc> units;
  std::vector<double> unit_theta;  // the double theta of each unit (result pose)
  std::vector<int> bin_tab;
  uint64_t total_bins = 0;
  uint64_t candidates = 0;
  int max_pts = 0;  // largest point count of any problem
};

// Size limits: violating one yields CGM_ERR_CAPACITY, never a wrong answer.
static const int kMaxPointsPerProblem = 16384;
static const double kMaxPointCells = 32000.0;  // |point| * inv_res must stay below this
static const uint64_t kMaxBins = 1ull << 28;
static const uint64_t kMaxCandidatesPerRegion = 0xFFFFFFFFull;

// _GridMap ctor arithmetic (gridmap.h:196-205).
GridGeom make_geom(float llx, float lly, float urx, float ury, float res, double kernel_range,
                   int kscale);
// _GridMap::world2grid / grid2world (gridmap.h:24-27, 45-48).
void world2grid(const GridGeom& g, float wx, float wy, int* ix, int* iy);
void grid2world(const GridGeom& g, int ix, int iy, float* wx, float* wy);
// ScanMatcher::initializeKernel (scan_matcher.cpp:38-61); column-major dim x dim.
std::vector<uint8_t> make_stamp(double resolution, double kernel_range, int kscale, int* dim);
// CharGrid::subsample (chargrid.cpp:98-122).
int subsample(const double* src_xy, int n, double res, double* dst_xy);

// Build the descriptor tables for n problems. pts/regions are packed back to back;
// returns CGM_OK or an error with *err set.
int build_plan(const GridGeom& g, int first_slot, int n, const double* pts_xy,
               const int* pts_counts, const float* regions, const int* region_counts,
               const SearchParams& p, SearchPlan* plan, std::string* err);

// Turn the survivors of one launch into the reference's mresvec per problem: per-chunk ordered
// maps with the "strictly better replaces" rule (chargrid.cpp:36-46), concatenation in chunk
// order (:298-303) and std::sort by score (:306-307).
void assemble_results(const GridGeom& g, const SearchPlan& plan, std::vector<Survivor>& survivors,
                      std::vector<std::vector<cgm_result> >* per_problem);

// The level ladder of hierarchicalSearch (chargrid.cpp:376-400).
std::vector<SearchParams> hierarchical_levels(const GridGeom& g, double theta_res,
                                              double max_score, double bin_x, double bin_y,
                                              double bin_theta, int n_levels);
// Next level's regions from a level's results (chargrid.cpp:322-331).
std::vector<float> refine_regions(const std::vector<cgm_result>& res, const SearchParams& level);

}  // namespace cgm
#endif
