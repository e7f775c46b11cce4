// Device back-end interface of the scan matcher. The product implementation is
// matcher_kernels.cu (CUDA, sm_100a). tests/hostsim/ holds a CPU stand-in of the same interface
// that exists ONLY so the host logic (planning, merging, C ABI) can be unit-tested on a box
// without a GPU; it is never linked into libcgmrslam_b200.so.
#ifndef CGM_MATCHER_DEVICE_H
#define CGM_MATCHER_DEVICE_H

#include <cstdint>
#include <string>
#include <vector>

#include "matcher_plan.h"

namespace cgm {

struct DeviceMatcher;  // opaque

int dev_device_count();
uint64_t dev_launch_count();

int dev_create(DeviceMatcher** out, int device, void* stream, int n_slots, const GridGeom& g,
               const uint8_t* stamp_colmajor, int stamp_dim, std::string* err);
void dev_destroy(DeviceMatcher* d);
void* dev_stream(const DeviceMatcher* d);

// Dense rows x cols copies (host layout has no row padding).
int dev_grid_download(DeviceMatcher* d, int slot, uint8_t* dst, std::string* err);
int dev_grid_upload(DeviceMatcher* d, int slot, const uint8_t* src, std::string* err);

// Replace the distance stamp (column-major dim x dim) / the value the reset path writes and the
// upper bound of any cell value the scoring kernels may assume; copy one slot's cells to a slot
// of another matcher of the same geometry on the same device.
int dev_set_stamp(DeviceMatcher* d, const uint8_t* stamp_colmajor, int stamp_dim, std::string* err);
void dev_set_bounds(DeviceMatcher* d, int fill_value, int max_cell);
int dev_copy_grid(DeviceMatcher* dst, int dst_slot, DeviceMatcher* src, int src_slot, std::string* err);

// Map building. stage = H2D of the packed points; launch = [reset] + stamp kernels (async).
// Passing n_points_total == 0 with reset == true stages a pure reset.
int dev_stage_map(DeviceMatcher* d, int first_slot, int n, const double* map_xy,
                  const int* counts, bool reset, std::string* err);
int dev_launch_map(DeviceMatcher* d, std::string* err);
void dev_clear_map_stage(DeviceMatcher* d);

// Search. stage = H2D of points + descriptor tables; launch = score + compact kernels (async);
// collect = sync + D2H of the survivors (and of the per-unit point counts k_theta, summed with
// the candidates of each unit into *cell_reads).
int dev_stage_search(DeviceMatcher* d, const SearchPlan& plan, const double* pts_xy,
                     int n_pts_total, std::string* err);
int dev_launch_search(DeviceMatcher* d, const SearchPlan& plan, int kernel_choice,
                      int* score_launches, std::string* err);
int dev_collect(DeviceMatcher* d, const SearchPlan& plan, std::vector<Survivor>* survivors,
                uint64_t* cell_reads, std::string* err);
int dev_sync(DeviceMatcher* d, std::string* err);
// Device time of the last launches, from events recorded on the matcher's stream around the
// kernels: ms[0] raster (0 if no map was launched since the last search), ms[1] scoring kernel,
// ms[2] compaction. Synchronises the stream.
int dev_last_timings(DeviceMatcher* d, float ms[3], std::string* err);

// Cold paths: sum of cells in [ax,bx) x [ay,by) clipped to the grid (countPoints), and cell
// values at given grid coordinates (-1 when outside; searchNonMatchedPoints).
int dev_window_sum(DeviceMatcher* d, int slot, int ax, int ay, int bx, int by, long long* sum,
                   std::string* err);
int dev_cells_at(DeviceMatcher* d, int slot, const int* gxy, int n, int* values,
                 std::string* err);

}  // namespace cgm
#endif
