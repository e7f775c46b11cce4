// score_tiled: production scoring kernel of the scan matcher (included by matcher_kernels.cu,
// inside namespace cgm). Computes exactly what score_global computes -- the loops of
// CharGrid::greedySearch, chargrid.cpp:239-287 -- for windows with a candidate stride of one cell.
//
// Idea. For one theta the score of offset (a, b) is  sum_p G[px_p + a][py_p + b]  over the k kept
// points. For a fixed point and a fixed row the offsets b, b+1, b+2, b+3 read four CONSECUTIVE
// bytes of one grid row (cell(x, y) = rows[x][y], gridmap.h:66-69). The kernel therefore
//   1. stages the part of the grid a job can touch (window + point extent) in shared memory,
//      zero-filled outside the grid (an out-of-grid cell adds 0 but still counts in k,
//      chargrid.cpp:272-276);
//   2. per theta rotates/quantises/dedups the points (chargrid.cpp:241-258) and sorts them into
//      four classes by the byte alignment of their column inside the staged tile;
//   3. per class, each thread owns ROWS consecutive rows x one aligned 32-bit word (four offsets,
//      shifted by the class) and issues ONE aligned 32-bit shared load per point and row. Up to
//      G = floor(255 / max_cell) words are added byte-wise before being widened into two packed
//      16-bit accumulators (0.5 ALU instruction per cell);
//   4. the four per-class partial sums are re-aligned through a small shared array, compared
//      against the integer acceptance threshold of this theta, and the rare survivors are merged
//      into the global per-bin arg-min table.
// The integer sums do not depend on summation order, so the result is bit-identical to the
// reference's sequential loop.

struct TiledJob {
  int region;   // index into the RegionDesc table
  int a0, ta;   // rows [a0, a0 + ta) of the region's window (offsets along x)
  int unit0, n_units;  // thetas [unit0, unit0 + n_units) (absolute unit indices)
  int rc;       // bound on |quantised point coordinate| for this problem, in cells
  int pad0, pad1;
};

struct TiledPlan {
  std::vector<TiledJob> jobs;
  int threads = 0;
  int group = 1;       // G
  size_t smem = 0;     // dynamic shared memory per CTA
  int max_pts = 0;
  std::string why_not;
};

const int kTiledRows = 4;  // ROWS

// Shared-memory layout of one job, derived identically on host (for sizing) and device.
struct TiledLayout {
  int nw;     // words per row group, padded to a multiple of 4 (bank mapping, see below)
  int n_rg;   // row groups = ceil(ta / ROWS)
  int fp;     // footprint pitch in bytes (multiple of 4)
  int fr;     // footprint rows
  int tw;     // words per row of one plane of the partial-sum array
  int off_tot, off_list, total;  // byte offsets
};

// Bank mapping: a warp's lanes are consecutive thread tiles tt = rg * nw + w and read word
// (rg * ROWS * fpw + w) + const. With ROWS * fpw == nw (mod 32) that is tt + const: conflict-free.
__host__ __device__ inline TiledLayout tiled_layout(int ny, int ta, int rc, int n_pts) {
  TiledLayout L;
  const int nw_real = (ny + 3 + 3) / 4;
  L.nw = (nw_real + 3) / 4 * 4;
  L.n_rg = (ta + kTiledRows - 1) / kTiledRows;
  int fpw = (2 * rc + 4) / 4 + 1 + L.nw;  // columns: point extent [-rc, rc] + alignment slack + window
  while ((kTiledRows * fpw - L.nw) % 32 != 0) ++fpw;
  L.fp = 4 * fpw;
  L.fr = L.n_rg * kTiledRows + 2 * rc + 1;
  L.tw = L.nw + 1;
  L.off_tot = L.fr * L.fp;
  L.off_list = L.off_tot + 4 * (L.n_rg * kTiledRows) * L.tw * 4;
  L.total = L.off_list + 2 * n_pts * 4 + 64;
  return L;
}

template <int G>
__global__ void __launch_bounds__(1024)
score_tiled(const uint8_t* __restrict__ grids, size_t slot_bytes, DevGeom g, int max_cell,
            const double* __restrict__ pts, const RegionDesc* __restrict__ regions,
            const ThetaDesc* __restrict__ units, const int* __restrict__ bin_tab, uint64_t* bins,
            double max_score, unsigned long long* scalars, const TiledJob* __restrict__ jobs) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int s_cnt[9];  // [0..3] class sizes, [4..7] fill cursors, [8] k
  __shared__ int s_thr;
  constexpr int R = kTiledRows;

  const TiledJob job = jobs[blockIdx.x];
  const RegionDesc reg = regions[job.region];
  const int rc = job.rc;
  const TiledLayout L = tiled_layout(reg.ny, job.ta, rc, reg.pts_n);
  uint8_t* F = smem;
  uint32_t* tot = reinterpret_cast<uint32_t*>(smem + L.off_tot);
  int* tmp = reinterpret_cast<int*>(smem + L.off_list);  // kept points, unordered: off | class
  int* lists = tmp + reg.pts_n;                           // the same, grouped by class
  const int plane = L.n_rg * R * L.tw;

  // ---- 1. stage the footprint ---------------------------------------------------------------
  // Tile origin in grid cells. Columns start at a multiple of 4 so that global loads are
  // aligned words; `shift` is what that costs in alignment slack.
  const int gx0 = reg.llx + job.a0 - rc;
  const int gy_lo = reg.lly - rc;
  const int gy0 = gy_lo & ~3;  // floor to a multiple of 4 (two's complement, also for negatives)
  const int shift = gy_lo - gy0;
  {
    const uint8_t* __restrict__ grid = grids + static_cast<size_t>(reg.slot) * slot_bytes;
    const int fpw = L.fp >> 2;
    for (int t = threadIdx.x; t < L.fr * fpw; t += blockDim.x) {
      const int r = t / fpw, cw = t - r * fpw;
      const int x = gx0 + r, y = gy0 + 4 * cw;
      uint32_t v = 0;
      if (x >= 0 && x < g.rows && y + 3 >= 0 && y < g.cols) {
        const uint8_t* row = grid + static_cast<size_t>(x) * g.pitch;
        if (y >= 0 && y + 3 < g.cols) {
          v = *reinterpret_cast<const uint32_t*>(row + y);
        } else {
#pragma unroll
          for (int bb = 0; bb < 4; ++bb)
            if (y + bb >= 0 && y + bb < g.cols) v |= static_cast<uint32_t>(row[y + bb]) << (8 * bb);
        }
      }
      reinterpret_cast<uint32_t*>(F)[t] = v;
    }
  }

  const double inv_res = static_cast<double>(g.inv_res);
  const double2* __restrict__ P = reinterpret_cast<const double2*>(pts) + reg.pts_off;
  const int n_tt = L.n_rg * L.nw;
  const int chunk_max = 65535 / max_cell;  // points per 16-bit accumulation run
  const uint32_t per_theta = static_cast<uint32_t>(reg.nx) * reg.ny;

  for (int uu = 0; uu < job.n_units; ++uu) {
    const int u = job.unit0 + uu;
    const ThetaDesc unit = units[u];
    if (threadIdx.x < 9) s_cnt[threadIdx.x] = 0;
    __syncthreads();  // also orders the footprint stores / the previous theta's final pass

    // ---- 2. rotate, quantise, dedup, classify ------------------------------------------------
    for (int i = threadIdx.x; i < reg.pts_n; i += blockDim.x) {
      const double2 p = P[i];
      const int2 ip = rotate_quantise(unit.c, unit.s, p.x, p.y, inv_res);
      int2 prev = make_int2(-10000, -10000);  // chargrid.cpp:242
      if (i > 0) {
        const double2 q = P[i - 1];
        prev = rotate_quantise(unit.c, unit.s, q.x, q.y, inv_res);
      }
      if (ip.x != prev.x || ip.y != prev.y) {
        if (ip.x < -rc || ip.x > rc || ip.y < -rc || ip.y > rc) {
          atomicExch(scalars + 2, 1ull);  // host-side bound violated: reported as an error
          continue;
        }
        const int cy = ip.y + rc + shift;  // column of offset b = 0 inside the tile
        const int s = cy & 3;
        atomicAdd(&s_cnt[s], 1);
        tmp[atomicAdd(&s_cnt[8], 1)] = ((ip.x + rc) * L.fp + (cy - s)) | s;  // fp, cy - s: multiples of 4
      }
    }
    __syncthreads();
    {
      const int c0 = s_cnt[0], c1 = s_cnt[1], c2 = s_cnt[2];
      const int kk = s_cnt[8];
      for (int i = threadIdx.x; i < kk; i += blockDim.x) {
        const int v = tmp[i];
        const int s = v & 3;
        const int first_of = (s > 0 ? c0 : 0) + (s > 1 ? c1 : 0) + (s > 2 ? c2 : 0);
        lists[first_of + atomicAdd(&s_cnt[4 + s], 1)] = v & ~3;
      }
    }
    __syncthreads();
    const int k = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
    if (threadIdx.x == 0) {
      s_thr = accept_threshold(k, max_cell * k, g.ikscale, max_score);
      atomicAdd(scalars, static_cast<unsigned long long>(k) *
                             (static_cast<unsigned long long>(job.ta) * reg.ny));
    }
    __syncthreads();
    const int thr = s_thr;
    if (thr == 0) continue;  // uniform: nothing can be accepted at this theta

    // ---- 3. accumulate, class by class ---------------------------------------------------------
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
      const int n_s = s_cnt[s];
      if (n_s == 0 && s > 0) continue;  // uniform
      const int* lst = lists + (s > 0 ? s_cnt[0] : 0) + (s > 1 ? s_cnt[1] : 0) + (s > 2 ? s_cnt[2] : 0);
      bool first = (s == 0);
      int done = 0;
      do {
        const int n_run = min(n_s - done, chunk_max);
        for (int tt = threadIdx.x; tt < n_tt; tt += blockDim.x) {
          const int rg = tt / L.nw, w = tt - rg * L.nw;
          const uint8_t* base = F + rg * R * L.fp + 4 * w;
          uint32_t acc_e[R], acc_o[R];
#pragma unroll
          for (int q = 0; q < R; ++q) acc_e[q] = acc_o[q] = 0;
          int p = 0;
          for (; p + G <= n_run; p += G) {
            int o[G];
#pragma unroll
            for (int gg = 0; gg < G; ++gg) o[gg] = lst[done + p + gg];
#pragma unroll
            for (int q = 0; q < R; ++q) {
              uint32_t v = *reinterpret_cast<const uint32_t*>(base + o[0] + q * L.fp);
#pragma unroll
              for (int gg = 1; gg < G; ++gg)
                v += *reinterpret_cast<const uint32_t*>(base + o[gg] + q * L.fp);
              acc_e[q] += v & 0x00FF00FFu;
              acc_o[q] += __byte_perm(v, 0, 0x4341);  // bytes 1 and 3 into the 16-bit lanes
            }
          }
          for (; p < n_run; ++p) {
            const int o1 = lst[done + p];
#pragma unroll
            for (int q = 0; q < R; ++q) {
              const uint32_t v = *reinterpret_cast<const uint32_t*>(base + o1 + q * L.fp);
              acc_e[q] += v & 0x00FF00FFu;
              acc_o[q] += __byte_perm(v, 0, 0x4341);
            }
          }
          // ---- 4a. re-align: byte t of word w is offset b = 4w + t - s, stored at c = b + 3,
          // plane c & 3, word c >> 2 (lanes with consecutive w hit consecutive words).
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const int row = rg * R + q;
            const uint32_t part[4] = {acc_e[q] & 0xFFFFu, acc_o[q] & 0xFFFFu, acc_e[q] >> 16,
                                      acc_o[q] >> 16};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int c = 4 * w + t - s + 3;
              uint32_t* dst = tot + (c & 3) * plane + row * L.tw + (c >> 2);
              if (first) *dst = part[t];
              else *dst += part[t];
            }
          }
        }
        done += n_run;
        first = false;
        __syncthreads();
      } while (done < n_s);
    }

    // ---- 4b. threshold, score, report ----------------------------------------------------------
    {
      const int n_c = job.ta * reg.ny;
      for (int t = threadIdx.x; t < n_c; t += blockDim.x) {
        const int row = t / reg.ny, b = t - row * reg.ny;
        const int c = b + 3;
        const int idsum = static_cast<int>(tot[(c & 3) * plane + row * L.tw + (c >> 2)]);
        if (idsum < thr) {
          const int a = job.a0 + row;
          const uint32_t cand =
              static_cast<uint32_t>(u - reg.theta_off) * per_theta + a * reg.ny + b;
          report(bins, reg, bin_tab, unit.bin_th, a, b, cand, score_of(idsum, k, g.ikscale));
        }
      }
    }
    // the __syncthreads at the top of the next iteration separates this pass from the next flush
  }
}

static cudaError_t tiled_configure(int smem_optin) {
  cudaError_t e;
  e = cudaFuncSetAttribute(score_tiled<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(score_tiled<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(score_tiled<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(score_tiled<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024);
  return e;
}

// Decide whether the staged search can run on score_tiled and cut it into jobs.
static void tiled_plan(const GridGeom& g, const SearchPlan& plan, const double* pts_xy,
                       int smem_optin, TiledPlan* out) {
  out->jobs.clear();
  out->why_not.clear();
  out->smem = 0;
  out->threads = 0;
  out->max_pts = plan.max_pts;
  if (plan.regions.empty() || plan.units.empty()) {
    out->why_not = "nothing to score";
    return;
  }
  if (g.max_cell <= 0) {
    out->why_not = "grid holds only zeros";
    return;
  }
  out->group = std::max(1, std::min(4, 255 / g.max_cell));
  const size_t budget = static_cast<size_t>(smem_optin) - 1024 - 64;
  size_t max_smem = 0;
  int max_tt = 0;
  size_t total_work = 0;
  // pass 1: per-region tile height
  struct Cut {
    int ta, rc;
  };
  std::vector<Cut> cuts(plan.regions.size());
  int last_problem_off = -1, rc_cached = 0;
  for (size_t r = 0; r < plan.regions.size(); ++r) {
    const RegionDesc& d = plan.regions[r];
    cuts[r].ta = 0;
    cuts[r].rc = 0;
    if (d.nx == 0 || d.ny == 0) continue;
    if (d.xs != 1 || d.ys != 1) {
      out->why_not = "candidate stride is not one cell";
      out->jobs.clear();
      return;
    }
    if (d.pts_off != last_problem_off) {
      double r2 = 0.0;
      for (int i = 0; i < d.pts_n; ++i) {
        const double x = pts_xy[2 * (d.pts_off + i)], y = pts_xy[2 * (d.pts_off + i) + 1];
        r2 = std::max(r2, x * x + y * y);
      }
      rc_cached = static_cast<int>(std::sqrt(r2) * static_cast<double>(g.inv_res)) + 2;
      last_problem_off = d.pts_off;
    }
    const int rc = rc_cached;
    int ta = d.nx;
    while (ta > 0 && static_cast<size_t>(tiled_layout(d.ny, ta, rc, d.pts_n).total) > budget)
      ta = (ta > 8) ? (ta + 1) / 2 : ta - 1;
    if (ta <= 0) {
      out->why_not = "window + point extent does not fit in shared memory";
      out->jobs.clear();
      return;
    }
    // balance the row tiles
    const int n_tiles = (d.nx + ta - 1) / ta;
    ta = (d.nx + n_tiles - 1) / n_tiles;
    cuts[r].ta = ta;
    cuts[r].rc = rc;
    total_work += static_cast<size_t>(n_tiles) * plan.regions_host[r].n_theta;
  }
  // pass 2: theta chunks so that there are enough jobs to fill the machine
  const size_t want_jobs = 148 * 4;
  for (size_t r = 0; r < plan.regions.size(); ++r) {
    const RegionDesc& d = plan.regions[r];
    if (cuts[r].ta == 0) continue;
    const int n_theta = plan.regions_host[r].n_theta;
    const int n_tiles = (d.nx + cuts[r].ta - 1) / cuts[r].ta;
    int per_job = n_theta;
    if (total_work < want_jobs * 8) per_job = 1;
    else if (total_work / n_theta < want_jobs)
      per_job = std::max<size_t>(1, std::min<size_t>(n_theta, total_work / want_jobs));
    for (int a0 = 0; a0 < d.nx; a0 += cuts[r].ta) {
      const int ta = std::min(cuts[r].ta, d.nx - a0);
      for (int t0 = 0; t0 < n_theta; t0 += per_job) {
        TiledJob j;
        j.region = static_cast<int>(r);
        j.a0 = a0;
        j.ta = ta;
        j.unit0 = d.theta_off + t0;
        j.n_units = std::min(per_job, n_theta - t0);
        j.rc = cuts[r].rc;
        j.pad0 = j.pad1 = 0;
        out->jobs.push_back(j);
      }
      const TiledLayout L = tiled_layout(d.ny, ta, cuts[r].rc, d.pts_n);
      max_smem = std::max(max_smem, static_cast<size_t>(L.total));
      max_tt = std::max(max_tt, L.n_rg * L.nw);
    }
    (void)n_tiles;
  }
  out->smem = max_smem;
  // one thread per thread tile where possible; otherwise the fewest equal passes
  const int passes = (max_tt + 1023) / 1024;
  out->threads = std::max(128, std::min(1024, ((max_tt + passes - 1) / passes + 31) / 32 * 32));
}

static cudaError_t tiled_launch(const uint8_t* grids, size_t slot_bytes, const DevGeom& dg,
                                int max_cell, const double* pts, const RegionDesc* regions,
                                const ThetaDesc* units, const int* bin_tab, uint64_t* bins,
                                double max_score, unsigned long long* scalars, const TiledJob* jobs,
                                const TiledPlan& tp, cudaStream_t stream) {
  const unsigned n = static_cast<unsigned>(tp.jobs.size());
  switch (tp.group) {
    case 1:
      score_tiled<1><<<n, tp.threads, tp.smem, stream>>>(grids, slot_bytes, dg, max_cell, pts,
                                                         regions, units, bin_tab, bins, max_score,
                                                         scalars, jobs);
      break;
    case 2:
      score_tiled<2><<<n, tp.threads, tp.smem, stream>>>(grids, slot_bytes, dg, max_cell, pts,
                                                         regions, units, bin_tab, bins, max_score,
                                                         scalars, jobs);
      break;
    case 3:
      score_tiled<3><<<n, tp.threads, tp.smem, stream>>>(grids, slot_bytes, dg, max_cell, pts,
                                                         regions, units, bin_tab, bins, max_score,
                                                         scalars, jobs);
      break;
    default:
      score_tiled<4><<<n, tp.threads, tp.smem, stream>>>(grids, slot_bytes, dg, max_cell, pts,
                                                         regions, units, bin_tab, bins, max_score,
                                                         scalars, jobs);
      break;
  }
  return cudaGetLastError();
}
