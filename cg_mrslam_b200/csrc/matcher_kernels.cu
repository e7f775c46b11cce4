// CUDA back-end of the scan matcher (sm_100a). Implements matcher_device.h.
//
// Kernels
//   raster_bands        resetGrid (scan_matcher.cpp:68-76) + addAndConvolvePoints/applyKernel
//                       (chargrid.h:205-216, chargrid.cpp:132-161): one CTA per band of whole grid
//                       rows, min-stamping in shared memory, the band out as one bulk async copy.
//   score_global        greedySearch inner loops (chargrid.cpp:239-287) straight from the
//                       L2/L1-resident grid; any stride, any point extent. Reference kernel.
//   score_tiled         the same arithmetic for stride-1 windows with the touched part of the
//                       grid staged in shared memory and four candidates per 32-bit load
//                       (DESIGN.md section 4). Production kernel.
//   compact_bins        drains the per-bin arg-min table into a dense survivor list.
//   window_sum, cells_at  cold paths (countPoints, searchNonMatchedPoints).
//
// Bit-exactness rules: every float/double operation that the reference performs is
// spelled with a round-to-nearest intrinsic so that nvcc cannot contract it into an FMA; cos/sin
// come from the host (ThetaDesc); integer sums are exact in any order.
#include <cuda_runtime.h>

#include <mutex>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "matcher_device.h"

namespace cgm {

namespace {

std::atomic<uint64_t> g_launches(0);

#define CGM_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      if (err) *err = std::string(#call) + ": " + cudaGetErrorString(e_);           \
      return CGM_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

// Geometry as the kernels need it.
struct DevGeom {
  float llx, lly, inv_res;
  int rows, cols, pitch;
  int fill_value;
  float ikscale;  // 1.f / (float)kscale, chargrid.cpp:260
};

// --------------------------------------------------------------------------------------------
// shared device helpers
// --------------------------------------------------------------------------------------------

// chargrid.cpp:249-250: rotate in double (two products, one sum, no FMA), scale by the float
// inverse resolution widened to double, truncate toward zero.
__device__ __forceinline__ int2 rotate_quantise(double c, double s, double x, double y,
                                                double inv_res) {
  const double px = __dsub_rn(__dmul_rn(c, x), __dmul_rn(s, y));
  const double py = __dadd_rn(__dmul_rn(s, x), __dmul_rn(c, y));
  int2 ip;
  ip.x = __double2int_rz(__dmul_rn(px, inv_res));
  ip.y = __double2int_rz(__dmul_rn(py, inv_res));
  return ip;
}

// chargrid.cpp:275-276: float(idsum) * ikscale, divided by k in double, narrowed to float.
__device__ __forceinline__ float score_of(int idsum, int k, float ikscale) {
  const float dsum = __fmul_rn(__int2float_rn(idsum), ikscale);
  return __double2float_rn(__ddiv_rn(static_cast<double>(dsum), static_cast<double>(k)));
}

// Smallest idsum that is NOT accepted by `score < maxScore` (chargrid.cpp:279). score_of is
// monotone in idsum for fixed k, so candidates are accepted iff idsum < threshold.
__device__ int accept_threshold(int k, int max_sum, float ikscale, double max_score) {
  if (k <= 0) return 0;  // dsum = maxScore + 1 is never accepted
  if (!(static_cast<double>(score_of(0, k, ikscale)) < max_score)) return 0;
  int lo = 0, hi = max_sum + 1;  // accepted(lo) is true; hi is beyond every reachable sum
  while (hi - lo > 1) {
    const int mid = lo + (hi - lo) / 2;
    if (static_cast<double>(score_of(mid, k, ikscale)) < max_score)
      lo = mid;
    else
      hi = mid;
  }
  return hi;
}

__device__ __forceinline__ void report(uint64_t* bins, const RegionDesc& reg, const int* bin_tab,
                                       int bin_th, int a, int b, uint32_t cand, float score) {
  const uint32_t entry =
      reg.bin_base +
      (static_cast<uint32_t>(bin_tab[reg.binx_off + a]) * reg.nby + bin_tab[reg.biny_off + b]) *
          reg.nbth +
      bin_th;
  const uint64_t key = (static_cast<uint64_t>(__float_as_uint(score)) << 32) | cand;
  atomicMin(reinterpret_cast<unsigned long long*>(bins + entry),
            static_cast<unsigned long long>(key));
}

// --------------------------------------------------------------------------------------------
// raster
// --------------------------------------------------------------------------------------------
const int kRasterThreads = 256, kRasterBatch = 1024;
const int kStampCells = 4096;  // a stamp of up to 63 x 63 cells (cgm_matcher_set_stamp checks)

// One CTA per band of `band_rows` whole grid rows of one slot. The band lives in shared memory as
// ints (min-stamping uses shared atomicMin), the slot's points are culled against the band's rows
// ONCE (a 32 x 128 tile per CTA meant every point was tested by 132 CTAs; a band makes that 30),
// and the finished band -- band_rows * pitch consecutive bytes of the slot -- leaves as ONE bulk
// asynchronous copy (cp.async.bulk shared -> global through the TMA unit) from a byte staging
// area. Shared memory: ints[band_rows][pitch] | bytes[band_rows][pitch].
__global__ void __launch_bounds__(kRasterThreads)
raster_bands(uint8_t* grids, size_t slot_bytes, DevGeom g, int first_slot, const double* pts,
             const int* pts_off, const uint8_t* stamp, int dim, int reset, int band_rows) {
  extern __shared__ __align__(128) uint8_t raster_smem[];
  __shared__ int2 list[kRasterBatch];
  __shared__ int n_list;
  __shared__ uchar4 cell_of[kStampCells];  // stamp cell -> (i, j, value): no divisions in the stamping loop
  int* tile = reinterpret_cast<int*>(raster_smem);
  const int pitch = g.pitch;
  uint8_t* out = raster_smem + static_cast<size_t>(band_rows) * pitch * sizeof(int);
  const int r0 = blockIdx.x * band_rows, nr = min(band_rows, g.rows - r0);
  const int slot = first_slot + blockIdx.y;
  uint8_t* grid = grids + static_cast<size_t>(slot) * slot_bytes;
  const int center = (dim - 1) / 2;
  const int cells = nr * pitch;
  const int dim2 = dim * dim;
  for (int t = threadIdx.x; t < dim2; t += kRasterThreads) {
    const int i = t / dim, j = t - i * dim;
    cell_of[t] = make_uchar4(static_cast<unsigned char>(i), static_cast<unsigned char>(j), stamp[j * dim + i], 0);  // ker[j*kRows+i]
  }

  if (reset) {
    const int v = static_cast<uint8_t>(g.fill_value);
    for (int t = threadIdx.x; t < cells; t += kRasterThreads) tile[t] = v;
  } else {  // stamp on top of what the slot holds: 4 cells per 32-bit load (pitch is a multiple of 16)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(grid + static_cast<size_t>(r0) * pitch);
    for (int t = threadIdx.x; t < cells / 4; t += kRasterThreads) {
      const uint32_t w = src[t];
      tile[4 * t] = w & 0xFF;
      tile[4 * t + 1] = (w >> 8) & 0xFF;
      tile[4 * t + 2] = (w >> 16) & 0xFF;
      tile[4 * t + 3] = w >> 24;
    }
  }
  const int p_begin = pts_off[blockIdx.y], p_end = pts_off[blockIdx.y + 1];
  for (int base = p_begin; base < p_end; base += kRasterBatch) {
    if (threadIdx.x == 0) n_list = 0;
    __syncthreads();
    for (int p = base + threadIdx.x; p < min(base + kRasterBatch, p_end); p += kRasterThreads) {
      // chargrid.h:209-212: double -> float, then world2grid (gridmap.h:24-27) in float + lrint.
      const float fx = __double2float_rn(pts[2 * p]), fy = __double2float_rn(pts[2 * p + 1]);
      const float gx = __fmul_rn(__fsub_rn(fx, g.llx), g.inv_res);
      const float gy = __fmul_rn(__fsub_rn(fy, g.lly), g.inv_res);
      if (!(fabsf(gx) < 1e9f) || !(fabsf(gy) < 1e9f)) continue;  // far outside: stamps nothing
      const int ix = __float2int_rn(gx), iy = __float2int_rn(gy);
      if (ix + center < r0 || ix - center >= r0 + nr || iy + center < 0 || iy - center >= g.cols) continue;
      list[atomicAdd(&n_list, 1)] = make_int2(ix, iy);
    }
    __syncthreads();
    // one warp per point, the lanes over the stamp's cells (chargrid.cpp:145-151)
    for (int p = threadIdx.x >> 5; p < n_list; p += kRasterThreads >> 5) {
      const int2 ip = list[p];
      for (int cell = threadIdx.x & 31; cell < dim2; cell += 32) {
        const uchar4 cv = cell_of[cell];
        const int r = ip.x + cv.x - center, c = ip.y + cv.y - center;
        if (r < r0 || r >= r0 + nr || c < 0 || c >= g.cols) continue;
        atomicMin(&tile[(r - r0) * pitch + c], static_cast<int>(cv.z));
      }
    }
    __syncthreads();
  }
  __syncthreads();
  // bytes, 4 cells per thread and store; the pad columns cols .. pitch hold 0
  for (int t = threadIdx.x; t < cells / 4; t += kRasterThreads) {
    const int c = (4 * t) % pitch;
    const int4 v = reinterpret_cast<const int4*>(tile)[t];  // one 16-byte load: no bank conflicts
    uint32_t w = 0;
    if (c < g.cols) w |= static_cast<uint32_t>(v.x & 0xFF);
    if (c + 1 < g.cols) w |= static_cast<uint32_t>(v.y & 0xFF) << 8;
    if (c + 2 < g.cols) w |= static_cast<uint32_t>(v.z & 0xFF) << 16;
    if (c + 3 < g.cols) w |= static_cast<uint32_t>(v.w & 0xFF) << 24;
    reinterpret_cast<uint32_t*>(out)[t] = w;
  }
  // make the generic-proxy writes visible to the async proxy, then one thread hands the band to TMA
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned src = static_cast<unsigned>(__cvta_generic_to_shared(out));
    uint8_t* dst = grid + static_cast<size_t>(r0) * pitch;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(cells)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the band is in global memory when the CTA retires
  }
}

// --------------------------------------------------------------------------------------------
// scoring, reference kernel: grid read through L1/L2
// --------------------------------------------------------------------------------------------
const int kGlobalThreads = 256, kGlobalChunk = 2048;

__global__ void __launch_bounds__(kGlobalThreads)
score_global(const uint8_t* __restrict__ grids, size_t slot_bytes, DevGeom g, int max_cell,
             const double* __restrict__ pts, const RegionDesc* __restrict__ regions,
             const ThetaDesc* __restrict__ units, const int* __restrict__ bin_tab, uint64_t* bins,
             double max_score, unsigned long long* cell_reads) {
  extern __shared__ short2 kept[];
  __shared__ int s_k, s_thr, s_box[4];
  const ThetaDesc unit = units[blockIdx.x];
  const RegionDesc reg = regions[unit.region];
  const int ncand = reg.nx * reg.ny;
  const int c_begin = blockIdx.y * kGlobalChunk;
  if (c_begin >= ncand) return;
  if (threadIdx.x == 0) {
    s_k = 0;
    s_box[0] = s_box[1] = INT_MAX;
    s_box[2] = s_box[3] = INT_MIN;
  }
  __syncthreads();
  const double inv_res = static_cast<double>(g.inv_res);
  const double2* P = reinterpret_cast<const double2*>(pts) + reg.pts_off;
  for (int i = threadIdx.x; i < reg.pts_n; i += kGlobalThreads) {
    const double2 p = P[i];
    const int2 ip = rotate_quantise(unit.c, unit.s, p.x, p.y, inv_res);
    int2 prev = make_int2(-10000, -10000);  // chargrid.cpp:242
    if (i > 0) {
      const double2 q = P[i - 1];
      prev = rotate_quantise(unit.c, unit.s, q.x, q.y, inv_res);
    }
    if (ip.x != prev.x || ip.y != prev.y) {  // :251; the order of the kept points is irrelevant
      kept[atomicAdd(&s_k, 1)] = make_short2(static_cast<short>(ip.x), static_cast<short>(ip.y));
      atomicMin(&s_box[0], ip.x);
      atomicMin(&s_box[1], ip.y);
      atomicMax(&s_box[2], ip.x);
      atomicMax(&s_box[3], ip.y);
    }
  }
  __syncthreads();
  const int k = s_k;
  if (threadIdx.x == 0) {
    s_thr = accept_threshold(k, max_cell * k, g.ikscale, max_score);
    if (blockIdx.y == 0)
      atomicAdd(cell_reads, static_cast<unsigned long long>(k) * static_cast<unsigned>(ncand));
  }
  __syncthreads();
  const int thr = s_thr;
  if (thr == 0) return;
  const uint8_t* __restrict__ grid = grids + static_cast<size_t>(reg.slot) * slot_bytes;
  // Whole footprint of this unit inside the grid => no per-cell bounds test.
  const bool all_inside =
      k > 0 && reg.llx + s_box[0] >= 0 && reg.lly + s_box[1] >= 0 &&
      static_cast<long long>(reg.llx) + static_cast<long long>(reg.nx - 1) * reg.xs + s_box[2] <
          g.rows &&
      static_cast<long long>(reg.lly) + static_cast<long long>(reg.ny - 1) * reg.ys + s_box[3] <
          g.cols;
  const int c_end = min(ncand, c_begin + kGlobalChunk);
  for (int c = c_begin + threadIdx.x; c < c_end; c += kGlobalThreads) {
    const int a = c / reg.ny, b = c - a * reg.ny;
    const int i = reg.llx + a * reg.xs, j = reg.lly + b * reg.ys;
    int idsum = 0;
    if (all_inside) {
      const uint8_t* base = grid + static_cast<size_t>(i) * g.pitch + j;
      for (int p = 0; p < k; ++p) {
        const short2 q = kept[p];
        idsum += base[static_cast<int>(q.x) * g.pitch + q.y];
      }
    } else {
      for (int p = 0; p < k; ++p) {
        const short2 q = kept[p];
        const int x = q.x + i, y = q.y + j;  // chargrid.cpp:270-273
        if (x >= 0 && y >= 0 && x < g.rows && y < g.cols)
          idsum += grid[static_cast<size_t>(x) * g.pitch + y];
      }
    }
    if (idsum < thr) {
      const uint32_t per_theta = static_cast<uint32_t>(ncand);
      const uint32_t cand = static_cast<uint32_t>(blockIdx.x - reg.theta_off) * per_theta + c;
      report(bins, reg, bin_tab, unit.bin_th, a, b, cand, score_of(idsum, k, g.ikscale));
    }
  }
}

// --------------------------------------------------------------------------------------------
// compaction of the bin table
// --------------------------------------------------------------------------------------------
__global__ void compact_bins(uint64_t* bins, uint64_t n_bins, Survivor* out, unsigned int* count) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n_bins;
       e += stride) {
    const uint64_t key = bins[e];
    if (key != kEmptyBin) {
      const unsigned int pos = atomicAdd(count, 1u);
      Survivor s;
      s.entry = static_cast<uint32_t>(e);
      s.pad = 0;
      s.key = key;
      out[pos] = s;
      bins[e] = kEmptyBin;
    }
  }
}

// --------------------------------------------------------------------------------------------
// cold paths
// --------------------------------------------------------------------------------------------
__global__ void window_sum(const uint8_t* grid, DevGeom g, int ax, int ay, int bx, int by,
                           unsigned long long* out) {
  // chargrid.cpp:425-432 restricted to the cells that exist.
  const int x0 = max(ax, 0), y0 = max(ay, 0), x1 = min(bx, g.rows), y1 = min(by, g.cols);
  unsigned long long local = 0;
  if (x1 > x0 && y1 > y0) {
    const long long w = y1 - y0, n = static_cast<long long>(x1 - x0) * w;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n;
         t += static_cast<long long>(gridDim.x) * blockDim.x)
      local += grid[static_cast<size_t>(x0 + t / w) * g.pitch + (y0 + t % w)];
  }
  for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

__global__ void cells_at(const uint8_t* grid, DevGeom g, const int* gxy, int n, int* values) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = gxy[2 * i], y = gxy[2 * i + 1];
  values[i] = (x >= 0 && y >= 0 && x < g.rows && y < g.cols)
                  ? grid[static_cast<size_t>(x) * g.pitch + y]
                  : -1;
}

// grow-only device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 4 + 64;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

#include "matcher_tiled.cuh"

struct DeviceMatcher {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int n_slots = 0;
  GridGeom geom;
  DevGeom dg;
  size_t slot_bytes = 0;
  uint8_t* grids = nullptr;
  uint8_t* stamp = nullptr;
  int stamp_dim = 0;
  int sm_count = 148;
  int smem_optin = 0;
  // map stage
  DevBuf<double> map_pts;
  DevBuf<int> map_off;
  int map_first = 0, map_n = 0;
  bool map_reset = false, map_valid = false;
  // search stage
  DevBuf<double> pts;
  DevBuf<RegionDesc> regions;
  DevBuf<ThetaDesc> units;
  DevBuf<int> bin_tab;
  DevBuf<uint64_t> bins;
  DevBuf<Survivor> surv;
  DevBuf<TiledJob> jobs;
  bool bins_dirty = false;
  unsigned int* surv_count = nullptr;     // device
  unsigned long long* scalars = nullptr;  // device: [0] cell_reads, [1] window sum
  // cold-path scratch
  DevBuf<int> scratch;
  // host-side tiling plan of the staged search
  TiledPlan tiled;
  // timing events: 0/1 around the raster kernels, 2/3 around the scoring kernel, 4 after compaction
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bool map_timed = false, search_timed = false;
};

int dev_device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

uint64_t dev_launch_count() { return g_launches.load(); }

static DevGeom make_dev_geom(const GridGeom& g) {
  DevGeom d;
  d.llx = g.llx;
  d.lly = g.lly;
  d.inv_res = g.inv_res;
  d.rows = g.rows;
  d.cols = g.cols;
  d.pitch = g.pitch;
  d.fill_value = g.fill_value;
  d.ikscale = static_cast<float>(1. / static_cast<float>(g.kscale));  // chargrid.cpp:260
  return d;
}

int dev_create(DeviceMatcher** out, int device, void* stream, int n_slots, const GridGeom& g,
               const uint8_t* stamp_colmajor, int stamp_dim, std::string* err) {
  *out = nullptr;
  int n_dev = 0;
  CGM_CUDA(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) {
    if (err) *err = "no such CUDA device";
    return CGM_ERR_CUDA;
  }
  CGM_CUDA(cudaSetDevice(device));
  DeviceMatcher* d = new DeviceMatcher();
  d->device = device;
  d->n_slots = n_slots;
  d->geom = g;
  d->dg = make_dev_geom(g);
  d->slot_bytes = static_cast<size_t>(g.rows) * g.pitch;
  d->stamp_dim = stamp_dim;
  cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&d->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaError_t e = cudaSuccess;
  if (stream) {
    d->stream = static_cast<cudaStream_t>(stream);
  } else {
    e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
    d->own_stream = e == cudaSuccess;
  }
  if (e == cudaSuccess)
    e = cudaMalloc(reinterpret_cast<void**>(&d->grids), d->slot_bytes * n_slots);
  if (e == cudaSuccess)
    e = cudaMalloc(reinterpret_cast<void**>(&d->stamp), static_cast<size_t>(stamp_dim) * stamp_dim);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d->surv_count), 16);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d->scalars), 64);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d->stamp, stamp_colmajor, static_cast<size_t>(stamp_dim) * stamp_dim,
                        cudaMemcpyHostToDevice, d->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d->grids, 0, d->slot_bytes * n_slots, d->stream);
  for (int i = 0; i < 5 && e == cudaSuccess; ++i) e = cudaEventCreate(&d->ev[i]);
  if (e == cudaSuccess) e = tiled_configure(d->smem_optin);
  if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
  if (e != cudaSuccess) {
    if (err) *err = std::string("dev_create: ") + cudaGetErrorString(e);
    dev_destroy(d);
    return e == cudaErrorMemoryAllocation ? CGM_ERR_ALLOC : CGM_ERR_CUDA;
  }
  *out = d;
  return CGM_OK;
}

void dev_destroy(DeviceMatcher* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  if (d->stream) cudaStreamSynchronize(d->stream);
  if (d->grids) cudaFree(d->grids);
  if (d->stamp) cudaFree(d->stamp);
  if (d->surv_count) cudaFree(d->surv_count);
  if (d->scalars) cudaFree(d->scalars);
  d->map_pts.release();
  d->map_off.release();
  d->pts.release();
  d->regions.release();
  d->units.release();
  d->bin_tab.release();
  d->bins.release();
  d->surv.release();
  d->jobs.release();
  d->scratch.release();
  for (int i = 0; i < 5; ++i)
    if (d->ev[i]) cudaEventDestroy(d->ev[i]);
  if (d->own_stream && d->stream) cudaStreamDestroy(d->stream);
  delete d;
}

void* dev_stream(const DeviceMatcher* d) { return d->stream; }

int dev_sync(DeviceMatcher* d, std::string* err) {
  CGM_CUDA(cudaSetDevice(d->device));
  CGM_CUDA(cudaStreamSynchronize(d->stream));
  return CGM_OK;
}

int dev_grid_download(DeviceMatcher* d, int slot, uint8_t* dst, std::string* err) {
  CGM_CUDA(cudaSetDevice(d->device));
  CGM_CUDA(cudaMemcpy2DAsync(dst, d->geom.cols, d->grids + slot * d->slot_bytes, d->geom.pitch,
                             d->geom.cols, d->geom.rows, cudaMemcpyDeviceToHost, d->stream));
  CGM_CUDA(cudaStreamSynchronize(d->stream));
  return CGM_OK;
}

int dev_grid_upload(DeviceMatcher* d, int slot, const uint8_t* src, std::string* err) {
  CGM_CUDA(cudaSetDevice(d->device));
  d->geom.max_cell = 255;  // arbitrary bytes may now be present
  CGM_CUDA(cudaMemsetAsync(d->grids + slot * d->slot_bytes, 0, d->slot_bytes, d->stream));
  CGM_CUDA(cudaMemcpy2DAsync(d->grids + slot * d->slot_bytes, d->geom.pitch, src, d->geom.cols,
                             d->geom.cols, d->geom.rows, cudaMemcpyHostToDevice, d->stream));
  CGM_CUDA(cudaStreamSynchronize(d->stream));
  return CGM_OK;
}

int dev_set_stamp(DeviceMatcher* d, const uint8_t* stamp_colmajor, int stamp_dim, std::string* err) {
  CGM_CUDA(cudaSetDevice(d->device));
  CGM_CUDA(cudaStreamSynchronize(d->stream));
  const size_t bytes = static_cast<size_t>(stamp_dim) * stamp_dim;
  if (stamp_dim != d->stamp_dim) {
    if (d->stamp) cudaFree(d->stamp);
    d->stamp = nullptr;
    d->stamp_dim = 0;
    CGM_CUDA(cudaMalloc(reinterpret_cast<void**>(&d->stamp), bytes ? bytes : 1));
    d->stamp_dim = stamp_dim;
  }
  CGM_CUDA(cudaMemcpyAsync(d->stamp, stamp_colmajor, bytes, cudaMemcpyHostToDevice, d->stream));
  CGM_CUDA(cudaStreamSynchronize(d->stream));
  return CGM_OK;
}

void dev_set_bounds(DeviceMatcher* d, int fill_value, int max_cell) {
  d->geom.fill_value = fill_value;
  d->geom.max_cell = max_cell;
  d->dg.fill_value = fill_value;
}

int dev_copy_grid(DeviceMatcher* dst, int dst_slot, DeviceMatcher* src, int src_slot, std::string* err) {
  if (dst->device != src->device || dst->slot_bytes != src->slot_bytes) {
    if (err) *err = "grid copy needs two matchers of one geometry on one device";
    return CGM_ERR_ARG;
  }
  CGM_CUDA(cudaSetDevice(dst->device));
  CGM_CUDA(cudaStreamSynchronize(src->stream));
  CGM_CUDA(cudaMemcpyAsync(dst->grids + dst_slot * dst->slot_bytes, src->grids + src_slot * src->slot_bytes,
                           src->slot_bytes, cudaMemcpyDeviceToDevice, dst->stream));
  CGM_CUDA(cudaStreamSynchronize(dst->stream));
  return CGM_OK;
}

int dev_stage_map(DeviceMatcher* d, int first_slot, int n, const double* map_xy, const int* counts,
                  bool reset, std::string* err) {
  CGM_CUDA(cudaSetDevice(d->device));
  std::vector<int> off(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    if (counts[i] < 0) {
      if (err) *err = "negative point count";
      return CGM_ERR_ARG;
    }
    off[i + 1] = off[i] + counts[i];
  }
  const int total = off[n];
  if (total && !map_xy) {
    if (err) *err = "null map points";
    return CGM_ERR_ARG;
  }
  CGM_CUDA(d->map_off.reserve(n + 1));
  CGM_CUDA(d->map_pts.reserve(2 * static_cast<size_t>(total) + 2));
  // pageable host memory: cudaMemcpyAsync stages synchronously, so `off` may die afterwards
  CGM_CUDA(cudaMemcpyAsync(d->map_off.p, off.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice,
                           d->stream));
  if (total)
    CGM_CUDA(cudaMemcpyAsync(d->map_pts.p, map_xy, 2 * static_cast<size_t>(total) * sizeof(double),
                             cudaMemcpyHostToDevice, d->stream));
  CGM_CUDA(cudaStreamSynchronize(d->stream));
  d->map_first = first_slot;
  d->map_n = n;
  d->map_reset = reset;
  d->map_valid = true;
  return CGM_OK;
}

int dev_launch_map(DeviceMatcher* d, std::string* err) {
  if (!d->map_valid) {
    if (err) *err = "no map staged";
    return CGM_ERR_ARG;
  }
  CGM_CUDA(cudaSetDevice(d->device));
  if (d->map_n == 0) return CGM_OK;
  // rows per band: as many as keep ints + bytes of a band within ~100 kB (two CTAs per SM)
  int band_rows = std::min(32, std::max(1, static_cast<int>(100 * 1024 / (5 * static_cast<size_t>(d->geom.pitch)))));
  if (band_rows > 8) band_rows &= ~7;
  const int tiles = (d->geom.rows + band_rows - 1) / band_rows;
  const size_t raster_smem_bytes = 5 * static_cast<size_t>(band_rows) * d->geom.pitch;
  {
    // the opt-in limit is a property of the FUNCTION (per device), shared by every matcher of the
    // process: raise it monotonically
    static std::mutex mu;
    static size_t limit[64] = {0};
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = limit[d->device & 63];
    if (raster_smem_bytes > 48 * 1024 && raster_smem_bytes > cur) {
      CGM_CUDA(cudaFuncSetAttribute(raster_bands, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(raster_smem_bytes)));
      cur = raster_smem_bytes;
    }
  }
  CGM_CUDA(cudaEventRecord(d->ev[0], d->stream));
  for (int s0 = 0; s0 < d->map_n; s0 += 65535) {
    const int ns = std::min(65535, d->map_n - s0);
    raster_bands<<<dim3(tiles, ns), kRasterThreads, raster_smem_bytes, d->stream>>>(
        d->grids, d->slot_bytes, d->dg, d->map_first + s0, d->map_pts.p, d->map_off.p + s0,
        d->stamp, d->stamp_dim, d->map_reset ? 1 : 0, band_rows);
    g_launches++;
  }
  CGM_CUDA(cudaGetLastError());
  CGM_CUDA(cudaEventRecord(d->ev[1], d->stream));
  d->map_timed = true;
  return CGM_OK;
}

void dev_clear_map_stage(DeviceMatcher* d) { d->map_valid = false; }

int dev_stage_search(DeviceMatcher* d, const SearchPlan& plan, const double* pts_xy,
                     int n_pts_total, std::string* err) {
  CGM_CUDA(cudaSetDevice(d->device));
  CGM_CUDA(d->pts.reserve(2 * static_cast<size_t>(n_pts_total) + 2));
  CGM_CUDA(d->regions.reserve(plan.regions.size()));
  CGM_CUDA(d->units.reserve(plan.units.size()));
  CGM_CUDA(d->bin_tab.reserve(plan.bin_tab.size() + 1));
  const size_t old_bins = d->bins.cap;
  CGM_CUDA(d->bins.reserve(plan.total_bins + 1));
  CGM_CUDA(d->surv.reserve(plan.total_bins + 1));
  if (d->bins.cap != old_bins || d->bins_dirty) {
    CGM_CUDA(cudaMemsetAsync(d->bins.p, 0xFF, d->bins.cap * sizeof(uint64_t), d->stream));
    d->bins_dirty = false;
  }
  if (n_pts_total)
    CGM_CUDA(cudaMemcpyAsync(d->pts.p, pts_xy, 2 * static_cast<size_t>(n_pts_total) * sizeof(double),
                             cudaMemcpyHostToDevice, d->stream));
  CGM_CUDA(cudaMemcpyAsync(d->regions.p, plan.regions.data(),
                           plan.regions.size() * sizeof(RegionDesc), cudaMemcpyHostToDevice,
                           d->stream));
  CGM_CUDA(cudaMemcpyAsync(d->units.p, plan.units.data(), plan.units.size() * sizeof(ThetaDesc),
                           cudaMemcpyHostToDevice, d->stream));
  if (!plan.bin_tab.empty())
    CGM_CUDA(cudaMemcpyAsync(d->bin_tab.p, plan.bin_tab.data(), plan.bin_tab.size() * sizeof(int),
                             cudaMemcpyHostToDevice, d->stream));
  // Tiling plan for the shared-memory kernel (empty when the geometry does not qualify).
  tiled_plan(d->geom, plan, pts_xy, d->smem_optin, &d->tiled);
  if (!d->tiled.jobs.empty()) {
    CGM_CUDA(d->jobs.reserve(d->tiled.jobs.size()));
    CGM_CUDA(cudaMemcpyAsync(d->jobs.p, d->tiled.jobs.data(),
                             d->tiled.jobs.size() * sizeof(TiledJob), cudaMemcpyHostToDevice,
                             d->stream));
  }
  CGM_CUDA(cudaStreamSynchronize(d->stream));
  return CGM_OK;
}

int dev_launch_search(DeviceMatcher* d, const SearchPlan& plan, int kernel_choice,
                      int* score_launches, std::string* err) {
  CGM_CUDA(cudaSetDevice(d->device));
  *score_launches = 0;
  CGM_CUDA(cudaMemsetAsync(d->surv_count, 0, sizeof(unsigned int), d->stream));
  CGM_CUDA(cudaMemsetAsync(d->scalars, 0, sizeof(unsigned long long), d->stream));
  CGM_CUDA(cudaMemsetAsync(d->scalars + 2, 0, sizeof(unsigned long long), d->stream));
  if (plan.units.empty()) return CGM_OK;
  bool use_tiled = !d->tiled.jobs.empty();
  if (kernel_choice == 1) use_tiled = false;
  if (kernel_choice == 2 && !use_tiled) {
    if (err) *err = "tiled kernel requested but the window does not qualify: " + d->tiled.why_not;
    return CGM_ERR_CAPACITY;
  }
  d->bins_dirty = true;
  CGM_CUDA(cudaEventRecord(d->ev[2], d->stream));
  if (use_tiled) {
    cudaError_t e = tiled_launch(d->grids, d->slot_bytes, d->dg, d->geom.max_cell, d->pts.p,
                                 d->regions.p, d->units.p, d->bin_tab.p, d->bins.p,
                                 plan.params.max_score, d->scalars, d->jobs.p, d->tiled, d->stream);
    if (e != cudaSuccess) {
      if (err) *err = std::string("score_tiled launch: ") + cudaGetErrorString(e);
      return CGM_ERR_CUDA;
    }
    g_launches++;
    *score_launches = 1;
  } else {
    int max_cand = 0;
    for (size_t r = 0; r < plan.regions.size(); ++r)
      max_cand = std::max(max_cand, plan.regions[r].nx * plan.regions[r].ny);
    const int chunks = (max_cand + kGlobalChunk - 1) / kGlobalChunk;
    const size_t smem = static_cast<size_t>(plan.max_pts) * sizeof(short2);
    if (smem > 48 * 1024)
      CGM_CUDA(cudaFuncSetAttribute(score_global, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
    const size_t n_units = plan.units.size();
    if (chunks > 65535) {
      if (err) *err = "window too large for one launch";
      return CGM_ERR_CAPACITY;
    }
    score_global<<<dim3(static_cast<unsigned>(n_units), chunks), kGlobalThreads, smem, d->stream>>>(
        d->grids, d->slot_bytes, d->dg, d->geom.max_cell, d->pts.p, d->regions.p, d->units.p,
        d->bin_tab.p, d->bins.p, plan.params.max_score, d->scalars);
    g_launches++;
    *score_launches = 1;
  }
  CGM_CUDA(cudaGetLastError());
  CGM_CUDA(cudaEventRecord(d->ev[3], d->stream));
  if (plan.total_bins) {
    const int blocks = static_cast<int>(
        std::min<uint64_t>((plan.total_bins + 255) / 256, static_cast<uint64_t>(d->sm_count) * 16));
    compact_bins<<<blocks, 256, 0, d->stream>>>(d->bins.p, plan.total_bins, d->surv.p,
                                                d->surv_count);
    g_launches++;
    CGM_CUDA(cudaGetLastError());
  }
  CGM_CUDA(cudaEventRecord(d->ev[4], d->stream));
  d->search_timed = true;
  return CGM_OK;
}

int dev_last_timings(DeviceMatcher* d, float ms[3], std::string* err) {
  CGM_CUDA(cudaSetDevice(d->device));
  CGM_CUDA(cudaStreamSynchronize(d->stream));
  ms[0] = ms[1] = ms[2] = 0.f;
  if (d->map_timed) CGM_CUDA(cudaEventElapsedTime(&ms[0], d->ev[0], d->ev[1]));
  if (d->search_timed) {
    CGM_CUDA(cudaEventElapsedTime(&ms[1], d->ev[2], d->ev[3]));
    CGM_CUDA(cudaEventElapsedTime(&ms[2], d->ev[3], d->ev[4]));
  }
  return CGM_OK;
}

int dev_collect(DeviceMatcher* d, const SearchPlan& plan, std::vector<Survivor>* survivors,
                uint64_t* cell_reads, std::string* err) {
  (void)plan;
  CGM_CUDA(cudaSetDevice(d->device));
  unsigned int n = 0;
  unsigned long long sc[3] = {0, 0, 0};
  CGM_CUDA(cudaMemcpyAsync(&n, d->surv_count, sizeof n, cudaMemcpyDeviceToHost, d->stream));
  CGM_CUDA(cudaMemcpyAsync(sc, d->scalars, sizeof sc, cudaMemcpyDeviceToHost, d->stream));
  CGM_CUDA(cudaStreamSynchronize(d->stream));
  d->bins_dirty = false;  // the compaction drained every bin it reported
  if (sc[2]) {
    if (err) *err = "score_tiled: a quantised point left the planned extent (internal bound)";
    return CGM_ERR_CAPACITY;
  }
  const unsigned long long reads = sc[0];
  survivors->resize(n);
  if (n) {
    CGM_CUDA(cudaMemcpyAsync(survivors->data(), d->surv.p, n * sizeof(Survivor),
                             cudaMemcpyDeviceToHost, d->stream));
    CGM_CUDA(cudaStreamSynchronize(d->stream));
  }
  *cell_reads = reads;
  return CGM_OK;
}

int dev_window_sum(DeviceMatcher* d, int slot, int ax, int ay, int bx, int by, long long* sum,
                   std::string* err) {
  CGM_CUDA(cudaSetDevice(d->device));
  CGM_CUDA(cudaMemsetAsync(d->scalars + 1, 0, sizeof(unsigned long long), d->stream));
  window_sum<<<d->sm_count, 256, 0, d->stream>>>(d->grids + slot * d->slot_bytes, d->dg, ax, ay, bx,
                                                 by, d->scalars + 1);
  g_launches++;
  unsigned long long v = 0;
  CGM_CUDA(cudaMemcpyAsync(&v, d->scalars + 1, sizeof v, cudaMemcpyDeviceToHost, d->stream));
  CGM_CUDA(cudaStreamSynchronize(d->stream));
  *sum = static_cast<long long>(v);
  return CGM_OK;
}

int dev_cells_at(DeviceMatcher* d, int slot, const int* gxy, int n, int* values, std::string* err) {
  CGM_CUDA(cudaSetDevice(d->device));
  CGM_CUDA(d->scratch.reserve(3 * static_cast<size_t>(n)));
  CGM_CUDA(cudaMemcpyAsync(d->scratch.p, gxy, 2 * static_cast<size_t>(n) * sizeof(int),
                           cudaMemcpyHostToDevice, d->stream));
  cells_at<<<(n + 255) / 256, 256, 0, d->stream>>>(d->grids + slot * d->slot_bytes, d->dg,
                                                   d->scratch.p, n, d->scratch.p + 2 * n);
  g_launches++;
  CGM_CUDA(cudaMemcpyAsync(values, d->scratch.p + 2 * n, n * sizeof(int), cudaMemcpyDeviceToHost,
                           d->stream));
  CGM_CUDA(cudaStreamSynchronize(d->stream));
  return CGM_OK;
}

}  // namespace cgm
