// The reference's sources chat on stdout (cout / printf). The drivers' RESULT lines therefore go to
// the file named by $CGM_OUT when it is set (tests), else to stdout. Include after every other header.
#ifndef CGM_TESTS_DRIVER_OUT_H
#define CGM_TESTS_DRIVER_OUT_H
#include <cstdio>
#include <cstdlib>
static inline FILE* cgm_out_file() {
  static FILE* f = std::getenv("CGM_OUT") ? std::fopen(std::getenv("CGM_OUT"), "w") : stdout;
  return f ? f : stdout;
}
#define printf(...) std::fprintf(cgm_out_file(), __VA_ARGS__)
#endif
