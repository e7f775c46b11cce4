// TEST INFRASTRUCTURE ONLY -- never linked into libcgmrslam_b200.so, never loaded by the package.
//
// A CPU stand-in for the device back-end interface (cg_mrslam_b200/csrc/matcher_device.h) so that
// the HOST logic of the product (matcher_plan.cpp: planning, bin tables, survivor decoding,
// per-chunk maps, sort; matcher_api.cpp: the C ABI) can be checked against the oracle on a box
// without a GPU (`pytest -m "not gpu"`). It plays the role of the kernels with plain loops and
// produces the same survivor stream the kernels produce. The GPU tests (`-m gpu`) exercise the
// real kernels through the real library; this file is not a fallback for anything.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "matcher_device.h"

namespace cgm {

struct DeviceMatcher {
  int n_slots;
  GridGeom g;
  std::vector<uint8_t> stamp;
  int stamp_dim;
  std::vector<std::vector<uint8_t> > grids;  // dense rows x cols
  // staged map
  std::vector<double> map_pts;
  std::vector<int> map_off;
  int map_first, map_n;
  bool map_reset, map_valid;
  // staged search
  std::vector<double> pts;
  std::vector<Survivor> survivors;
  uint64_t cell_reads;
};

static uint64_t g_launches = 0;

int dev_device_count() { return 1; }
uint64_t dev_launch_count() { return g_launches; }

int dev_create(DeviceMatcher** out, int, void*, int n_slots, const GridGeom& g,
               const uint8_t* stamp, int dim, std::string*) {
  DeviceMatcher* d = new DeviceMatcher();
  d->n_slots = n_slots;
  d->g = g;
  d->stamp.assign(stamp, stamp + dim * dim);
  d->stamp_dim = dim;
  d->grids.assign(n_slots, std::vector<uint8_t>(static_cast<size_t>(g.rows) * g.cols, 0));
  d->map_valid = false;
  d->cell_reads = 0;
  *out = d;
  return CGM_OK;
}
void dev_destroy(DeviceMatcher* d) { delete d; }
void* dev_stream(const DeviceMatcher*) { return nullptr; }
int dev_sync(DeviceMatcher*, std::string*) { return CGM_OK; }
int dev_last_timings(DeviceMatcher*, float ms[3], std::string*) {
  ms[0] = ms[1] = ms[2] = 0.f;
  return CGM_OK;
}

int dev_grid_download(DeviceMatcher* d, int slot, uint8_t* dst, std::string*) {
  std::memcpy(dst, d->grids[slot].data(), d->grids[slot].size());
  return CGM_OK;
}
int dev_grid_upload(DeviceMatcher* d, int slot, const uint8_t* src, std::string*) {
  std::memcpy(d->grids[slot].data(), src, d->grids[slot].size());
  d->g.max_cell = 255;
  return CGM_OK;
}

int dev_set_stamp(DeviceMatcher* d, const uint8_t* stamp, int dim, std::string*) {
  d->stamp.assign(stamp, stamp + static_cast<size_t>(dim) * dim);
  d->stamp_dim = dim;
  return CGM_OK;
}

void dev_set_bounds(DeviceMatcher* d, int fill_value, int max_cell) {
  d->g.fill_value = fill_value;
  d->g.max_cell = max_cell;
}

int dev_copy_grid(DeviceMatcher* dst, int dst_slot, DeviceMatcher* src, int src_slot, std::string*) {
  if (dst->grids[dst_slot].size() != src->grids[src_slot].size()) return CGM_ERR_ARG;
  dst->grids[dst_slot] = src->grids[src_slot];
  return CGM_OK;
}

int dev_stage_map(DeviceMatcher* d, int first_slot, int n, const double* xy, const int* counts,
                  bool reset, std::string*) {
  d->map_off.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) d->map_off[i + 1] = d->map_off[i] + counts[i];
  d->map_pts.assign(xy, xy + 2 * static_cast<size_t>(d->map_off[n]));
  d->map_first = first_slot;
  d->map_n = n;
  d->map_reset = reset;
  d->map_valid = true;
  return CGM_OK;
}

int dev_launch_map(DeviceMatcher* d, std::string*) {
  const GridGeom& g = d->g;
  const int dim = d->stamp_dim, center = (dim - 1) / 2;
  for (int s = 0; s < d->map_n; ++s) {
    std::vector<uint8_t>& grid = d->grids[d->map_first + s];
    if (d->map_reset) std::fill(grid.begin(), grid.end(), static_cast<uint8_t>(g.fill_value));
    for (int p = d->map_off[s]; p < d->map_off[s + 1]; ++p) {
      int ix, iy;
      world2grid(g, static_cast<float>(d->map_pts[2 * p]), static_cast<float>(d->map_pts[2 * p + 1]),
                 &ix, &iy);
      for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j) {
          const int r = ix + i - center, c = iy + j - center;
          if (r < 0 || c < 0 || r >= g.rows || c >= g.cols) continue;
          uint8_t& v = grid[static_cast<size_t>(r) * g.cols + c];
          v = std::min(v, d->stamp[j * dim + i]);
        }
    }
    ++g_launches;
  }
  return CGM_OK;
}
void dev_clear_map_stage(DeviceMatcher* d) { d->map_valid = false; }

int dev_stage_search(DeviceMatcher* d, const SearchPlan&, const double* pts_xy, int n,
                     std::string*) {
  d->pts.assign(pts_xy, pts_xy + 2 * static_cast<size_t>(n));
  return CGM_OK;
}

int dev_launch_search(DeviceMatcher* d, const SearchPlan& plan, int, int* launches, std::string*) {
  const GridGeom& g = d->g;
  std::map<uint32_t, uint64_t> bins;
  d->cell_reads = 0;
  const float ikscale = static_cast<float>(1. / static_cast<float>(g.kscale));
  for (size_t u = 0; u < plan.units.size(); ++u) {
    const ThetaDesc& t = plan.units[u];
    const RegionDesc& r = plan.regions[t.region];
    const std::vector<uint8_t>& grid = d->grids[r.slot];
    std::vector<int> ipx, ipy;
    int px = -10000, py = -10000;
    for (int i = 0; i < r.pts_n; ++i) {
      const double x = d->pts[2 * (r.pts_off + i)], y = d->pts[2 * (r.pts_off + i) + 1];
      const double rx = t.c * x - t.s * y, ry = t.s * x + t.c * y;
      const int ix = static_cast<int>(rx * g.inv_res), iy = static_cast<int>(ry * g.inv_res);
      if (ix != px || iy != py) {
        ipx.push_back(ix);
        ipy.push_back(iy);
        px = ix;
        py = iy;
      }
    }
    const int k = static_cast<int>(ipx.size());
    d->cell_reads += static_cast<uint64_t>(k) * r.nx * r.ny;
    for (int a = 0; a < r.nx; ++a)
      for (int b = 0; b < r.ny; ++b) {
        const int i = r.llx + a * r.xs, j = r.lly + b * r.ys;
        int idsum = 0;
        for (int p = 0; p < k; ++p) {
          const int x = ipx[p] + i, y = ipy[p] + j;
          if (x >= 0 && y >= 0 && x < g.rows && y < g.cols)
            idsum += grid[static_cast<size_t>(x) * g.cols + y];
        }
        float dsum = static_cast<float>(idsum) * ikscale;
        dsum = k ? static_cast<float>(dsum / static_cast<double>(k))
                 : static_cast<float>(plan.params.max_score + 1);
        if (!(dsum < plan.params.max_score)) continue;
        uint32_t bits;
        std::memcpy(&bits, &dsum, 4);
        const uint32_t cand = static_cast<uint32_t>(u - r.theta_off) * (r.nx * r.ny) + a * r.ny + b;
        const uint64_t key = (static_cast<uint64_t>(bits) << 32) | cand;
        const uint32_t entry =
            r.bin_base +
            (static_cast<uint32_t>(plan.bin_tab[r.binx_off + a]) * r.nby + plan.bin_tab[r.biny_off + b]) *
                r.nbth +
            t.bin_th;
        std::map<uint32_t, uint64_t>::iterator it = bins.find(entry);
        if (it == bins.end()) bins[entry] = key;
        else if (key < it->second) it->second = key;
      }
  }
  d->survivors.clear();
  for (std::map<uint32_t, uint64_t>::reverse_iterator it = bins.rbegin(); it != bins.rend(); ++it) {
    Survivor s;  // deliberately not in entry order: the kernels report in arbitrary order too
    s.entry = it->first;
    s.pad = 0;
    s.key = it->second;
    d->survivors.push_back(s);
  }
  *launches = 1;
  g_launches += 2;
  return CGM_OK;
}

int dev_collect(DeviceMatcher* d, const SearchPlan&, std::vector<Survivor>* out, uint64_t* reads,
                std::string*) {
  *out = d->survivors;
  *reads = d->cell_reads;
  return CGM_OK;
}

int dev_window_sum(DeviceMatcher* d, int slot, int ax, int ay, int bx, int by, long long* sum,
                   std::string*) {
  const GridGeom& g = d->g;
  long long s = 0;
  for (int i = std::max(ax, 0); i < std::min(bx, g.rows); ++i)
    for (int j = std::max(ay, 0); j < std::min(by, g.cols); ++j)
      s += d->grids[slot][static_cast<size_t>(i) * g.cols + j];
  *sum = s;
  return CGM_OK;
}

int dev_cells_at(DeviceMatcher* d, int slot, const int* gxy, int n, int* values, std::string*) {
  const GridGeom& g = d->g;
  for (int i = 0; i < n; ++i) {
    const int x = gxy[2 * i], y = gxy[2 * i + 1];
    values[i] = (x >= 0 && y >= 0 && x < g.rows && y < g.cols)
                    ? d->grids[slot][static_cast<size_t>(x) * g.cols + y]
                    : -1;
  }
  return CGM_OK;
}

}  // namespace cgm
