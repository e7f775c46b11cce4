// TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the reference's correlative scan matcher.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load the library built from this file; the product (cg_mrslam_b200/) never does.
//
// Parity status: PINNED. tests/test_oracle_matcher.py checks every function here bit-for-bit
// against the reference's own chargrid.cpp compiled verbatim (oracle/_ref/libref_chargrid.so,
// recipe in oracle/Makefile) and against the committed golden vectors in tests/golden/.
//
// Each function cites the reference lines it restates. Arithmetic types are kept exactly:
// float where the reference uses float, double where it uses double, truncation vs lrint,
// no FMA contraction (build with -ffp-contract=off).
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

namespace {

struct OGrid {
  float llx, lly, urx, ury;  // gridmap.h:185-186
  float res, inv_res;        // gridmap.h:182-183
  int rows, cols;            // rows = x extent, cols = y extent (gridmap.h:134)
  int kscale;                // chargrid.cpp:126
  std::vector<uint8_t> cells;  // row-major [x][y]: cell(x,y)=rows[x][y], gridmap.h:66-69
  uint8_t& at(int x, int y) { return cells[static_cast<size_t>(x) * cols + y]; }
  bool inside(int x, int y) const { return x >= 0 && y >= 0 && x < rows && y < cols; }  // gridmap.h:55-58
};

// gridmap.h:24-27: float subtraction, float multiply, lrint (current rounding mode = nearest even).
inline void world2grid(const OGrid& g, float wx, float wy, int* ix, int* iy) {
  *ix = static_cast<int>(lrint((wx - g.llx) * g.inv_res));
  *iy = static_cast<int>(lrint((wy - g.lly) * g.inv_res));
}

// gridmap.h:45-48: float multiply then float add.
inline void grid2world(const OGrid& g, int ix, int iy, float* wx, float* wy) {
  *wx = g.llx + (g.res * static_cast<float>(ix));
  *wy = g.lly + (g.res * static_cast<float>(iy));
}

struct OResult {
  double x, y, th, score;
};

// chargrid.h:68-85: bin coordinates are truncated ints stored as double, ordered (ix, iy, ith).
struct BinKey {
  double ix, iy, ith;
  bool operator<(const BinKey& o) const {
    if (ix < o.ix) return true;
    if (ix == o.ix && iy < o.iy) return true;
    if (ix == o.ix && iy == o.iy && ith < o.ith) return true;
    return false;
  }
};

inline BinKey bin_of(const OResult& r, double dx, double dy, double dth) {
  BinKey k;
  k.ix = static_cast<int>(r.x / dx);
  k.iy = static_cast<int>(r.y / dy);
  k.ith = static_cast<int>(r.th / dth);
  return k;
}

// chargrid.cpp:36-46: keep the strictly better score per bin (first found wins ties).
inline void add_pruned(std::map<BinKey, OResult>& m, const OResult& r, double dx, double dy,
                       double dth) {
  BinKey k = bin_of(r, dx, dy, dth);
  std::map<BinKey, OResult>::iterator it = m.find(k);
  if (it != m.end()) {
    if (it->second.score > r.score) it->second = r;
  } else {
    m.insert(std::make_pair(k, r));
  }
}

struct ScoreLess {  // chargrid.h:62-66
  bool operator()(const OResult& a, const OResult& b) const { return a.score < b.score; }
};

// chargrid.cpp:208-308. The OpenMP region is executed chunk by chunk (same chunking, same
// per-chunk maps, same concatenation order), which yields the same mresvec sequence.
void greedy_search(const OGrid& g, const double* pts, int n, const float* regions, int nreg,
                   double step_x, double step_y, double theta_res, double max_score, double dx,
                   double dy, double dth, std::vector<OResult>* out) {
  int x_steps = static_cast<int>(step_x / g.res);  // double / float -> double -> int (:214-215)
  int y_steps = static_cast<int>(step_y / g.res);
  if (x_steps <= 0) x_steps = 1;
  if (y_steps <= 0) y_steps = 1;

  size_t n_threads = nreg < 4 ? static_cast<size_t>(nreg) : 4;  // :223-224
  out->clear();
  if (n_threads == 0) return;  // the reference divides by zero here (:225); callers never do this
  size_t chunk = static_cast<size_t>(nreg) / n_threads;
  std::vector<std::map<BinKey, OResult> > maps(n_threads);
  std::vector<int> ipx(n), ipy(n);

  for (size_t tid = 0; tid < n_threads; ++tid) {
    size_t rmin = tid * chunk;
    size_t rmax = (tid == n_threads - 1) ? static_cast<size_t>(nreg) : (tid + 1) * chunk;
    for (size_t r = rmin; r < rmax; ++r) {
      const float* reg = regions + 6 * r;
      int llx, lly, urx, ury;
      world2grid(g, reg[0], reg[1], &llx, &lly);  // :237-238
      world2grid(g, reg[3], reg[4], &urx, &ury);
      for (double t = reg[2]; t < reg[5]; t += theta_res) {  // :239, float bounds widened
        double c = cos(t), s = sin(t);
        int prev_x = -10000, prev_y = -10000, k = 0;
        for (int i = 0; i < n; ++i) {  // :246-258
          double px = c * pts[2 * i] - s * pts[2 * i + 1];
          double py = s * pts[2 * i] + c * pts[2 * i + 1];
          int ix = static_cast<int>(px * g.inv_res);  // double * float -> double -> trunc
          int iy = static_cast<int>(py * g.inv_res);
          if (ix != prev_x || iy != prev_y) {
            ipx[k] = ix;
            ipy[k] = iy;
            ++k;
            prev_x = ix;
            prev_y = iy;
          }
        }
        float ikscale = 1. / static_cast<float>(g.kscale);  // :260
        for (int i = llx; i < urx; i += x_steps) {
          for (int j = lly; j < ury; j += y_steps) {
            int idsum = 0;
            for (int p = 0; p < k; ++p) {  // :268-274
              int cx = ipx[p] + i, cy = ipy[p] + j;
              if (g.inside(cx, cy)) idsum += g.cells[static_cast<size_t>(cx) * g.cols + cy];
            }
            float dsum = static_cast<float>(idsum) * ikscale;                     // :275
            dsum = k ? (dsum / static_cast<double>(k)) : max_score + 1;           // :276
            float wx, wy;
            grid2world(g, i, j, &wx, &wy);                                        // :277
            if (dsum < max_score) {                                               // :279
              OResult mr = {wx, wy, t, dsum};
              add_pruned(maps[tid], mr, dx, dy, dth);
            }
          }
        }
      }
    }
  }
  for (size_t tid = 0; tid < n_threads; ++tid)  // :298-303
    for (std::map<BinKey, OResult>::iterator it = maps[tid].begin(); it != maps[tid].end(); ++it)
      out->push_back(it->second);
  std::sort(out->begin(), out->end(), ScoreLess());  // :306-307 (same libstdc++ algorithm)
}

// chargrid.cpp:310-344 with the parameter ladder of :376-400.
void hierarchical_search(const OGrid& g, const double* pts, int n, const float* regions, int nreg,
                         double theta_res, double max_score, double dx, double dy, double dth,
                         int n_levels, std::vector<OResult>* out) {
  struct Level {
    double sx, sy, sth, bx, by, bth;
  };
  std::vector<Level> levels;
  for (int i = n_levels - 1; i >= 0; --i) {  // :383-393
    int m = static_cast<int>(pow(2, i));
    int mtheta = (m / 2 < 1) ? m : m / 2;
    Level l = {m * g.res, m * g.res, mtheta * theta_res, dx * m, dy * m, dth * m};
    levels.push_back(l);
  }
  out->clear();
  std::vector<float> cur(regions, regions + 6 * static_cast<size_t>(nreg));
  for (size_t li = 0; li + 1 < levels.size(); ++li) {  // :315-334
    const Level& l = levels[li];
    greedy_search(g, pts, n, cur.data(), static_cast<int>(cur.size() / 6), l.sx, l.sy, l.sth,
                  max_score, l.bx, l.by, l.bth, out);
    if (out->empty()) break;
    cur.clear();
    for (size_t i = 0; i < out->size(); ++i) {  // :322-331, double math then narrowed to float
      const OResult& b = (*out)[i];
      double lo[3] = {-(l.bx * .5) + b.x, -(l.by * .5) + b.y, -(l.bth * .5) + b.th};
      double hi[3] = {(l.bx * .5) + b.x, (l.by * .5) + b.y, (l.bth * .5) + b.th};
      for (int c = 0; c < 3; ++c) cur.push_back(static_cast<float>(lo[c]));
      for (int c = 0; c < 3; ++c) cur.push_back(static_cast<float>(hi[c]));
    }
  }
  if (!out->empty()) {  // :336-343
    const Level& l = levels.back();
    greedy_search(g, pts, n, cur.data(), static_cast<int>(cur.size() / 6), l.sx, l.sy, l.sth,
                  max_score, l.bx, l.by, l.bth, out);
  }
}

int emit(const std::vector<OResult>& res, double* out4, int cap) {
  int n = static_cast<int>(res.size());
  for (int i = 0; i < n && i < cap; ++i) {
    out4[4 * i + 0] = res[i].x;
    out4[4 * i + 1] = res[i].y;
    out4[4 * i + 2] = res[i].th;
    out4[4 * i + 3] = res[i].score;
  }
  return n;
}

}  // namespace

extern "C" {

// _GridMap ctor, gridmap.h:196-214: size = trunc((ur - ll) * invRes) in float.
void* orc_grid_create(float llx, float lly, float urx, float ury, float res, int kscale) {
  OGrid* g = new OGrid;
  g->llx = llx;
  g->lly = lly;
  g->urx = urx;
  g->ury = ury;
  g->res = res;
  g->inv_res = 1. / g->res;  // double divide narrowed to float (gridmap.h:202)
  float sx = (urx - llx) * g->inv_res, sy = (ury - lly) * g->inv_res;
  g->rows = static_cast<int>(sx);
  g->cols = static_cast<int>(sy);
  if (g->rows <= 0 || g->cols <= 0) g->rows = g->cols = 0;
  g->kscale = kscale;
  g->cells.assign(static_cast<size_t>(g->rows) * g->cols, 0);
  return g;
}

void orc_grid_destroy(void* h) { delete static_cast<OGrid*>(h); }

void orc_grid_size(void* h, int* rows, int* cols) {
  *rows = static_cast<OGrid*>(h)->rows;
  *cols = static_cast<OGrid*>(h)->cols;
}

// ScanMatcher::resetGrid, scan_matcher.cpp:68-76 (value narrowed to unsigned char).
void orc_grid_fill(void* h, int value) {
  OGrid* g = static_cast<OGrid*>(h);
  std::fill(g->cells.begin(), g->cells.end(), static_cast<uint8_t>(value));
}

// ScanMatcher::initializeKernel, scan_matcher.cpp:38-61. Column-major dim x dim stamp,
// dim = 2*int(range/res)+1. Returns dim; writes at most cap bytes.
int orc_make_stamp(double resolution, double kernel_range, int kscale, unsigned char* dst,
                   int cap) {
  int size = static_cast<int>(kernel_range / resolution);
  int dim = 2 * size + 1;
  if (dim * dim > cap) return -dim;
  int K1 = static_cast<int>(resolution * kscale);
  int K2 = static_cast<int>(kernel_range * kscale);
  std::vector<unsigned char> k(static_cast<size_t>(dim) * dim,
                               static_cast<unsigned char>(static_cast<char>(K2)));
  for (int j = 0; j <= size; ++j) {
    for (int i = 0; i <= size; ++i) {
      char distance = static_cast<char>(K1 * sqrt(static_cast<double>(j * j + i * i)));
      if (distance > K2) continue;
      unsigned char d = static_cast<unsigned char>(distance);
      k[static_cast<size_t>(size + j) * dim + (size + i)] = d;  // (row, col) -> col*dim + row
      k[static_cast<size_t>(size + j) * dim + (size - i)] = d;
      k[static_cast<size_t>(size - j) * dim + (size + i)] = d;
      k[static_cast<size_t>(size - j) * dim + (size - i)] = d;
    }
  }
  std::memcpy(dst, k.data(), k.size());
  return dim;
}

// addAndConvolvePoints (chargrid.h:205-216) + applyKernel (chargrid.cpp:132-161).
void orc_grid_raster(void* h, const double* xy, int n, const unsigned char* stamp_colmajor,
                     int kdim) {
  OGrid* g = static_cast<OGrid*>(h);
  int center = (kdim - 1) / 2;
  for (int p = 0; p < n; ++p) {
    float fx = static_cast<float>(xy[2 * p]), fy = static_cast<float>(xy[2 * p + 1]);
    int r, c;
    world2grid(*g, fx, fy, &r, &c);
    for (int i = 0; i < kdim; ++i) {
      int io = r + i - center;
      if (io < 0 || io >= g->rows) continue;
      for (int j = 0; j < kdim; ++j) {
        int jo = c + j - center;
        if (jo < 0 || jo >= g->cols) continue;
        uint8_t& v = g->at(io, jo);
        uint8_t kv = stamp_colmajor[j * kdim + i];
        v = (kv < v) ? kv : v;
      }
    }
  }
}

void orc_grid_download(void* h, unsigned char* dst) {
  OGrid* g = static_cast<OGrid*>(h);
  if (!g->cells.empty()) std::memcpy(dst, g->cells.data(), g->cells.size());
}

void orc_grid_upload(void* h, const unsigned char* src) {
  OGrid* g = static_cast<OGrid*>(h);
  if (!g->cells.empty()) std::memcpy(g->cells.data(), src, g->cells.size());
}

void orc_world2grid(void* h, float x, float y, int* ix, int* iy) {
  world2grid(*static_cast<OGrid*>(h), x, y, ix, iy);
}

void orc_grid2world(void* h, int ix, int iy, float* x, float* y) {
  grid2world(*static_cast<OGrid*>(h), ix, iy, x, y);
}

// CharGrid::subsample, chargrid.cpp:98-122: buckets keyed by truncated (x/res, y/res), visited in
// (ix, iy) order; bucket mean = running sum (input order) * (1/count).
int orc_subsample(const double* xy, int n, double res, double* dst_xy) {
  double ires = 1. / res;
  struct Acc {
    double sx, sy;
    int cnt;
  };
  std::map<std::pair<int, int>, Acc> acc;
  for (int i = 0; i < n; ++i) {
    std::pair<int, int> key(static_cast<int>(ires * xy[2 * i]), static_cast<int>(ires * xy[2 * i + 1]));
    std::map<std::pair<int, int>, Acc>::iterator it = acc.find(key);
    if (it == acc.end()) {
      Acc a = {0.0 + xy[2 * i], 0.0 + xy[2 * i + 1], 1};
      acc.insert(std::make_pair(key, a));
    } else {
      it->second.sx += xy[2 * i];
      it->second.sy += xy[2 * i + 1];
      it->second.cnt++;
    }
  }
  int k = 0;
  for (std::map<std::pair<int, int>, Acc>::iterator it = acc.begin(); it != acc.end(); ++it, ++k) {
    double w = 1. / static_cast<double>(it->second.cnt);
    dst_xy[2 * k] = it->second.sx * w;
    dst_xy[2 * k + 1] = it->second.sy * w;
  }
  return k;
}

int orc_greedy_search(void* h, const double* xy, int n, const float* regions, int nreg,
                      double step_x, double step_y, double step_th, double max_score, double dx,
                      double dy, double dth, double* out4, int cap) {
  std::vector<OResult> res;
  greedy_search(*static_cast<OGrid*>(h), xy, n, regions, nreg, step_x, step_y, step_th, max_score,
                dx, dy, dth, &res);
  return emit(res, out4, cap);
}

// chargrid.cpp:196-206: search step = (resolution, resolution, thetaRes), resolution is float.
int orc_greedy_search_res(void* h, const double* xy, int n, const float* regions, int nreg,
                          double theta_res, double max_score, double dx, double dy, double dth,
                          double* out4, int cap) {
  OGrid* g = static_cast<OGrid*>(h);
  std::vector<OResult> res;
  greedy_search(*g, xy, n, regions, nreg, g->res, g->res, theta_res, max_score, dx, dy, dth, &res);
  return emit(res, out4, cap);
}

int orc_hierarchical_search(void* h, const double* xy, int n, const float* regions, int nreg,
                            double theta_res, double max_score, double dx, double dy, double dth,
                            int n_levels, double* out4, int cap) {
  std::vector<OResult> res;
  hierarchical_search(*static_cast<OGrid*>(h), xy, n, regions, nreg, theta_res, max_score, dx, dy,
                      dth, n_levels, &res);
  return emit(res, out4, cap);
}

// CharGrid::countPoints, chargrid.cpp:417-441.
double orc_count_points(void* h, float llx, float lly, float urx, float ury) {
  OGrid* g = static_cast<OGrid*>(h);
  int ax, ay, bx, by;
  world2grid(*g, llx, lly, &ax, &ay);
  world2grid(*g, urx, ury, &bx, &by);
  int isum = 0;
  for (int i = ax; i < bx; ++i)
    for (int j = ay; j < by; ++j)
      if (g->inside(i, j)) isum += g->cells[static_cast<size_t>(i) * g->cols + j];
  int visited = (bx - ax) * (by - ay);
  return static_cast<float>(isum) / static_cast<float>(visited);
}

// CharGrid::searchNonMatchedPoints, chargrid.cpp:444-455.
int orc_search_non_matched(void* h, const double* xy, int n, double max_score, double* dst_xy) {
  OGrid* g = static_cast<OGrid*>(h);
  float ikscale = 1. / static_cast<float>(g->kscale);
  int k = 0;
  for (int i = 0; i < n; ++i) {
    int gx, gy;
    world2grid(*g, static_cast<float>(xy[2 * i]), static_cast<float>(xy[2 * i + 1]), &gx, &gy);
    if (g->inside(gx, gy)) {
      double value = static_cast<float>(g->cells[static_cast<size_t>(gx) * g->cols + gy]) * ikscale;
      if (value > max_score) {
        dst_xy[2 * k] = xy[2 * i];
        dst_xy[2 * k + 1] = xy[2 * i + 1];
        ++k;
      }
    }
  }
  return k;
}

}  // extern "C"
