// Host-side planning and result assembly of the scan matcher. See matcher_plan.h.
// Build with -ffp-contract=off: the float/double expressions below must round exactly like the
// reference's (no fused multiply-add).
#include "matcher_plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <utility>

namespace cgm {

GridGeom make_geom(float llx, float lly, float urx, float ury, float res, double kernel_range,
                   int kscale) {
  GridGeom g;
  g.llx = llx;
  g.lly = lly;
  g.urx = urx;
  g.ury = ury;
  g.res = res;
  g.inv_res = static_cast<float>(1. / g.res);  // gridmap.h:202: double quotient narrowed to float
  // gridmap.h:203-204: (upperRight - lowerLeft) * inverseResolution in float, truncated to int.
  float ex = (urx - llx) * g.inv_res;
  float ey = (ury - lly) * g.inv_res;
  g.rows = static_cast<int>(ex);
  g.cols = static_cast<int>(ey);
  if (g.rows <= 0 || g.cols <= 0) g.rows = g.cols = 0;
  g.pitch = (g.cols + 15) / 16 * 16;
  g.kscale = kscale;
  g.fill_value = static_cast<int>(kernel_range * kscale);  // scan_matcher.cpp:69
  g.max_cell = static_cast<uint8_t>(g.fill_value);
  return g;
}

void world2grid(const GridGeom& g, float wx, float wy, int* ix, int* iy) {
  // gridmap.h:26: float difference, float product, lrint (round half to even).
  float fx = (wx - g.llx) * g.inv_res;
  float fy = (wy - g.lly) * g.inv_res;
  *ix = static_cast<int>(lrint(fx));
  *iy = static_cast<int>(lrint(fy));
}

void grid2world(const GridGeom& g, int ix, int iy, float* wx, float* wy) {
  // gridmap.h:47: float product, then float sum.
  float px = g.res * static_cast<float>(ix);
  float py = g.res * static_cast<float>(iy);
  *wx = g.llx + px;
  *wy = g.lly + py;
}

std::vector<uint8_t> make_stamp(double resolution, double kernel_range, int kscale, int* dim_out) {
  // scan_matcher.cpp:39-47. The stamp is stored column-major like the reference's MatrixXChar;
  // it is symmetric, so the storage order only matters for readers of cgm_matcher_stamp().
  const int half = static_cast<int>(kernel_range / resolution);
  const int dim = 2 * half + 1;
  const int k1 = static_cast<int>(resolution * kscale);
  const int k2 = static_cast<int>(kernel_range * kscale);
  std::vector<uint8_t> stamp(static_cast<size_t>(dim) * dim,
                             static_cast<uint8_t>(static_cast<char>(k2)));
  for (int a = 0; a <= half; ++a) {
    for (int b = 0; b <= half; ++b) {
      // scan_matcher.cpp:51-53: the distance is narrowed to (signed) char before the comparison.
      const char d = static_cast<char>(k1 * std::sqrt(static_cast<double>(a * a + b * b)));
      if (d > k2) continue;
      const uint8_t v = static_cast<uint8_t>(d);
      const int lo_a = half - a, hi_a = half + a, lo_b = half - b, hi_b = half + b;
      stamp[static_cast<size_t>(hi_a) * dim + hi_b] = v;
      stamp[static_cast<size_t>(hi_a) * dim + lo_b] = v;
      stamp[static_cast<size_t>(lo_a) * dim + hi_b] = v;
      stamp[static_cast<size_t>(lo_a) * dim + lo_b] = v;
    }
  }
  *dim_out = dim;
  return stamp;
}

int subsample(const double* src, int n, double res, double* dst) {
  // chargrid.cpp:98-122: bucket key = truncated (x/res, y/res); buckets are emitted in ascending
  // (ix, iy) order; the mean is (running sum in input order) * (1 / count).
  const double ires = 1. / res;
  struct Bucket {
    double sx, sy;
    int count;
  };
  typedef std::pair<int, int> Cell;
  std::map<Cell, Bucket> buckets;
  for (int i = 0; i < n; ++i) {
    const double x = src[2 * i], y = src[2 * i + 1];
    const Cell key(static_cast<int>(ires * x), static_cast<int>(ires * y));
    std::map<Cell, Bucket>::iterator it = buckets.find(key);
    if (it == buckets.end()) {
      Bucket b = {0.0, 0.0, 0};
      it = buckets.insert(std::make_pair(key, b)).first;
    }
    it->second.sx += x;
    it->second.sy += y;
    it->second.count += 1;
  }
  int k = 0;
  for (std::map<Cell, Bucket>::const_iterator it = buckets.begin(); it != buckets.end(); ++it) {
    const double w = 1. / static_cast<double>(it->second.count);
    dst[2 * k] = it->second.sx * w;
    dst[2 * k + 1] = it->second.sy * w;
    ++k;
  }
  return k;
}

namespace {

// DiscreteTriplet (chargrid.h:68-85): truncation toward zero of value / bin width.
inline int bin_coord(double value, double width) { return static_cast<int>(value / width); }

}  // namespace

int build_plan(const GridGeom& g, int first_slot, int n, const double* pts_xy,
               const int* pts_counts, const float* regions, const int* region_counts,
               const SearchParams& p, SearchPlan* plan, std::string* err) {
  *plan = SearchPlan();
  plan->params = p;
  if (!(p.step_theta > 0.0)) {
    *err = "theta step must be positive (the reference would loop forever, chargrid.cpp:239)";
    return CGM_ERR_ARG;
  }
  if (!(p.bin_x > 0.0) || !(p.bin_y > 0.0) || !(p.bin_theta > 0.0)) {
    *err = "result discretisation must be positive";
    return CGM_ERR_ARG;
  }
  // chargrid.cpp:214-221
  int xs = static_cast<int>(p.step_x / g.res);
  int ys = static_cast<int>(p.step_y / g.res);
  if (xs <= 0) xs = 1;
  if (ys <= 0) ys = 1;

  plan->n_problems = n;
  int pts_off = 0, reg_off = 0;
  uint64_t bin_base = 0;
  for (int pr = 0; pr < n; ++pr) {
    const int np = pts_counts[pr];
    const int nr = region_counts[pr];
    if (np < 0 || nr < 0) {
      *err = "negative count";
      return CGM_ERR_ARG;
    }
    if (np > kMaxPointsPerProblem) {
      *err = "too many scan points in one problem";
      return CGM_ERR_CAPACITY;
    }
    for (int i = 0; i < np; ++i) {
      const double x = pts_xy[2 * (pts_off + i)], y = pts_xy[2 * (pts_off + i) + 1];
      const double r2 = x * x + y * y;
      const double lim = kMaxPointCells / g.inv_res;
      if (!(r2 < lim * lim)) {
        *err = "scan point too far from the sensor for 16-bit cell offsets";
        return CGM_ERR_CAPACITY;
      }
    }
    plan->max_pts = std::max(plan->max_pts, np);
    plan->problem_slot.push_back(first_slot + pr);
    plan->problem_first_region.push_back(static_cast<int>(plan->regions.size()));
    plan->problem_n_regions.push_back(nr);
    // chargrid.cpp:223-225
    const int n_chunks = nr < 4 ? nr : 4;
    plan->problem_n_chunks.push_back(n_chunks);
    const int chunk_size = n_chunks ? nr / n_chunks : 0;

    for (int r = 0; r < nr; ++r) {
      const float* reg = regions + 6 * static_cast<size_t>(reg_off + r);
      RegionDesc d;
      std::memset(&d, 0, sizeof d);
      RegionHost h;
      h.problem = pr;
      h.chunk = (chunk_size == 0) ? 0 : std::min(r / chunk_size, n_chunks - 1);
      d.slot = first_slot + pr;
      d.pts_off = pts_off;
      d.pts_n = np;
      int urx, ury;
      world2grid(g, reg[0], reg[1], &d.llx, &d.lly);  // chargrid.cpp:237-238
      world2grid(g, reg[3], reg[4], &urx, &ury);
      d.xs = xs;
      d.ys = ys;
      d.nx = urx > d.llx ? static_cast<int>((static_cast<int64_t>(urx) - d.llx + xs - 1) / xs) : 0;
      d.ny = ury > d.lly ? static_cast<int>((static_cast<int64_t>(ury) - d.lly + ys - 1) / ys) : 0;
      d.theta_off = static_cast<int>(plan->units.size());

      // theta ladder, chargrid.cpp:239: repeated double addition from the float lower bound.
      std::vector<double> ladder;
      const double t_hi = reg[5];
      for (double t = reg[2]; t < t_hi; t += p.step_theta) {
        ladder.push_back(t);
        if (ladder.size() > (1u << 22)) {
          *err = "more than 4M theta steps in one region";
          return CGM_ERR_CAPACITY;
        }
      }
      const bool empty = d.nx == 0 || d.ny == 0 || np == 0 || ladder.empty();
      if (empty) {
        // No candidate of this region can be accepted (no offsets, or k = 0 => score =
        // maxScore + 1, chargrid.cpp:276). Keep the descriptor so indices stay aligned.
        d.nx = d.ny = 0;
        h.n_theta = 0;
        h.n_bins = 0;
        d.bin_base = static_cast<uint32_t>(bin_base);
        plan->regions.push_back(d);
        plan->regions_host.push_back(h);
        continue;
      }
      if (static_cast<uint64_t>(d.nx) * d.ny * ladder.size() > kMaxCandidatesPerRegion) {
        *err = "more than 2^32 candidates in one region";
        return CGM_ERR_CAPACITY;
      }

      // Result-bin coordinates per i step / j step / theta step (chargrid.cpp:277-284 with
      // DiscreteTriplet chargrid.h:69-73): pose = (float grid2world widened to double, t).
      std::vector<int> bx(d.nx), by(d.ny), bt(ladder.size());
      for (int a = 0; a < d.nx; ++a) {
        float wx, wy;
        grid2world(g, d.llx + a * xs, 0, &wx, &wy);
        bx[a] = bin_coord(static_cast<double>(wx), p.bin_x);
      }
      for (int b = 0; b < d.ny; ++b) {
        float wx, wy;
        grid2world(g, 0, d.lly + b * ys, &wx, &wy);
        by[b] = bin_coord(static_cast<double>(wy), p.bin_y);
      }
      for (size_t k = 0; k < ladder.size(); ++k) bt[k] = bin_coord(ladder[k], p.bin_theta);
      const int bx0 = *std::min_element(bx.begin(), bx.end());
      const int by0 = *std::min_element(by.begin(), by.end());
      const int bt0 = *std::min_element(bt.begin(), bt.end());
      const int64_t nbx = static_cast<int64_t>(*std::max_element(bx.begin(), bx.end())) - bx0 + 1;
      const int64_t nby = static_cast<int64_t>(*std::max_element(by.begin(), by.end())) - by0 + 1;
      const int64_t nbt = static_cast<int64_t>(*std::max_element(bt.begin(), bt.end())) - bt0 + 1;
      if (nbx > (1 << 20) || nby > (1 << 20) || nbt > (1 << 20) ||
          static_cast<uint64_t>(nbx) * nby * nbt > kMaxBins ||
          bin_base + static_cast<uint64_t>(nbx) * nby * nbt > kMaxBins) {
        *err = "result-bin table too large (discretisation too fine for the window)";
        return CGM_ERR_CAPACITY;
      }
      d.nby = static_cast<int>(nby);
      d.nbth = static_cast<int>(nbt);
      d.binx_off = static_cast<int>(plan->bin_tab.size());
      for (int a = 0; a < d.nx; ++a) plan->bin_tab.push_back(bx[a] - bx0);
      d.biny_off = static_cast<int>(plan->bin_tab.size());
      for (int b = 0; b < d.ny; ++b) plan->bin_tab.push_back(by[b] - by0);
      d.bin_base = static_cast<uint32_t>(bin_base);
      h.n_theta = static_cast<int>(ladder.size());
      h.n_bins = static_cast<uint32_t>(nbx * nby * nbt);
      bin_base += h.n_bins;

      const int region_index = static_cast<int>(plan->regions.size());
      for (size_t k = 0; k < ladder.size(); ++k) {
        ThetaDesc u;
        u.c = std::cos(ladder[k]);  // chargrid.cpp:241
        u.s = std::sin(ladder[k]);
        u.region = region_index;
        u.bin_th = bt[k] - bt0;
        plan->units.push_back(u);
        plan->unit_theta.push_back(ladder[k]);
      }
      plan->candidates += static_cast<uint64_t>(d.nx) * d.ny * ladder.size();
      plan->regions.push_back(d);
      plan->regions_host.push_back(h);
    }
    pts_off += np;
    reg_off += nr;
  }
  plan->total_bins = bin_base;
  return CGM_OK;
}

namespace {

struct BinKey {  // DiscreteTriplet ordering, chargrid.h:75-83
  int ix, iy, ith;
  bool operator<(const BinKey& o) const {
    if (ix != o.ix) return ix < o.ix;
    if (iy != o.iy) return iy < o.iy;
    return ith < o.ith;
  }
};

struct ByEntry {
  bool operator()(const Survivor& a, const Survivor& b) const { return a.entry < b.entry; }
};

struct ByScore {  // MatcherResultScoreComparator, chargrid.h:62-66
  bool operator()(const cgm_result& a, const cgm_result& b) const { return a.score < b.score; }
};

}  // namespace

void assemble_results(const GridGeom& g, const SearchPlan& plan, std::vector<Survivor>& survivors,
                      std::vector<std::vector<cgm_result> >* per_problem) {
  per_problem->assign(plan.n_problems, std::vector<cgm_result>());
  // Entry order == (region, bin) order; regions of a problem are consecutive.
  std::sort(survivors.begin(), survivors.end(), ByEntry());

  // first bin entry of every region, for entry -> region lookup
  std::vector<uint64_t> base(plan.regions.size() + 1);
  for (size_t r = 0; r < plan.regions.size(); ++r) base[r] = plan.regions[r].bin_base;
  base[plan.regions.size()] = plan.total_bins;

  size_t cursor = 0;
  for (int pr = 0; pr < plan.n_problems; ++pr) {
    const int r0 = plan.problem_first_region[pr];
    const int r1 = r0 + plan.problem_n_regions[pr];
    const int n_chunks = plan.problem_n_chunks[pr];
    std::vector<std::map<BinKey, cgm_result> > maps(n_chunks);
    for (int r = r0; r < r1; ++r) {
      const RegionDesc& d = plan.regions[r];
      const RegionHost& h = plan.regions_host[r];
      const uint64_t end = base[r] + h.n_bins;
      while (cursor < survivors.size() && survivors[cursor].entry < end) {
        const Survivor& s = survivors[cursor++];
        const uint32_t cand = static_cast<uint32_t>(s.key & 0xFFFFFFFFull);
        const uint32_t bits = static_cast<uint32_t>(s.key >> 32);
        float score;
        std::memcpy(&score, &bits, sizeof score);
        const uint32_t per_theta = static_cast<uint32_t>(d.nx) * d.ny;
        const uint32_t ti = cand / per_theta;
        const uint32_t rem = cand % per_theta;
        const int a = static_cast<int>(rem / d.ny), b = static_cast<int>(rem % d.ny);
        float wx, wy;
        grid2world(g, d.llx + a * d.xs, d.lly + b * d.ys, &wx, &wy);  // chargrid.cpp:277
        cgm_result res;
        res.x = wx;
        res.y = wy;
        res.theta = plan.unit_theta[d.theta_off + ti];
        res.score = score;
        BinKey key;
        key.ix = bin_coord(res.x, plan.params.bin_x);
        key.iy = bin_coord(res.y, plan.params.bin_y);
        key.ith = bin_coord(res.theta, plan.params.bin_theta);
        // addToPrunedMap, chargrid.cpp:36-46. Regions are visited in order, so an equal score
        // from a later region never displaces the earlier one.
        std::map<BinKey, cgm_result>& m = maps[h.chunk];
        std::map<BinKey, cgm_result>::iterator it = m.find(key);
        if (it == m.end()) {
          m.insert(std::make_pair(key, res));
        } else if (it->second.score > res.score) {
          it->second = res;
        }
      }
    }
    std::vector<cgm_result>& out = (*per_problem)[pr];
    for (int c = 0; c < n_chunks; ++c)  // chargrid.cpp:298-303
      for (std::map<BinKey, cgm_result>::const_iterator it = maps[c].begin(); it != maps[c].end();
           ++it)
        out.push_back(it->second);
    std::sort(out.begin(), out.end(), ByScore());  // chargrid.cpp:306-307
  }
}

std::vector<SearchParams> hierarchical_levels(const GridGeom& g, double theta_res,
                                              double max_score, double bin_x, double bin_y,
                                              double bin_theta, int n_levels) {
  std::vector<SearchParams> levels;
  for (int i = n_levels - 1; i >= 0; --i) {  // chargrid.cpp:383-393
    const int m = static_cast<int>(std::pow(2, i));
    const int mtheta = (m / 2 < 1) ? m : m / 2;
    SearchParams p;
    const float cell = m * g.res;  // int * float stays float (chargrid.cpp:388)
    p.step_x = cell;
    p.step_y = cell;
    p.step_theta = mtheta * theta_res;
    p.max_score = max_score;
    p.bin_x = bin_x * m;
    p.bin_y = bin_y * m;
    p.bin_theta = bin_theta * m;
    levels.push_back(p);
  }
  return levels;
}

std::vector<float> refine_regions(const std::vector<cgm_result>& res, const SearchParams& level) {
  std::vector<float> out;
  out.reserve(res.size() * 6);
  for (size_t i = 0; i < res.size(); ++i) {  // chargrid.cpp:322-331
    const double hx = level.bin_x * .5, hy = level.bin_y * .5, ht = level.bin_theta * .5;
    const double lo[3] = {-hx + res[i].x, -hy + res[i].y, -ht + res[i].theta};
    const double hi[3] = {hx + res[i].x, hy + res[i].y, ht + res[i].theta};
    for (int c = 0; c < 3; ++c) out.push_back(static_cast<float>(lo[c]));
    for (int c = 0; c < 3; ++c) out.push_back(static_cast<float>(hi[c]));
  }
  return out;
}

}  // namespace cgm
