// The CUDA device new CharGrid / SparseOptimizer objects are created on. One process per GPU
// normally selects it with CUDA_VISIBLE_DEVICES; CGM_DEVICE overrides the default of 0, and a host
// program may assign cgm::default_device() before it constructs its objects.
#ifndef CGM_DEVICE_HPP
#define CGM_DEVICE_HPP
#include <cstdlib>
namespace cgm {
inline int& default_device() {
  static int device = std::getenv("CGM_DEVICE") ? std::atoi(std::getenv("CGM_DEVICE")) : 0;
  return device;
}
}  // namespace cgm
#endif
