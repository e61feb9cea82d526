// mcb_host.hpp -- host-side helpers shared by the two C-ABI translation units
// (mcb_layer.cu: one layer; mcb_world.cu: a whole multi-GPU run).
#pragma once
#include <string>

#include "../../include/mcb200.h"
#include "mcb_kernels.cuh"

namespace mcb {

// record the message mcb200_last_error() returns (thread-local) and hand back `code`
int fail(int code, const std::string &msg);

#define MCB_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return mcb::fail(e__ == cudaErrorMemoryAllocation ? MCB200_ERR_NOMEM : MCB200_ERR_CUDA, \
                       std::string(#expr) + ": " + cudaGetErrorString(e__));                  \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// per-cell event constants, src/layer.cpp:131-133 (host code is built with -ffp-contract=off
// and no -mfma, like the reference: plain IEEE float ops)
CellXs host_cell_xs(float sig, float absorption_rate);

// 128-bit two's-complement accumulator (4 little-endian 32-bit digits, LSB 2^kAccLsbLog2)
// -> double, rounded once
double acc_to_double(const unsigned d[kAccDigits]);

// the device layout of a tally (64-bit halves [2][ncell]) -> cell-major 32-bit digits
// out[4*c + j] for the first m cells
void acc_halves_to_digits(const unsigned *raw, int ncell, int m, unsigned *out_4m);

// the 32-bit Philox key of a run from the 64-bit seed the caller gave (src/layer.cpp:36)
inline unsigned philox_key(unsigned long long seed) { return (unsigned)(seed ^ (seed >> 32)); }

const char *last_error_cstr();
void set_last_error(const std::string &m);
std::string get_last_error();

}  // namespace mcb
