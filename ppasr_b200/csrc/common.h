// Shared host-side helpers: error reporting for the C-ABI and CUDA call checking.
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace ppasr {

// Thread-local last error message returned by ppasr_b200_last_error().
void set_last_error(const std::string& msg);
const char* get_last_error();

int device_sm_count();
bool pdl_enabled();
void set_pdl_enabled(bool on);

// number of kernels launched by this library (all contexts); read by bench.py for "gpu_launches"
void count_launch();
long long launch_count();
void add_launches(long long n);  // kernels replayed by a CUDA graph (counted once at capture, n per replay)

}  // namespace ppasr

#define PPASR_OK 0
#define PPASR_ERR_INVALID 1
#define PPASR_ERR_CUDA 2
#define PPASR_ERR_STATE 3

#define PPASR_CUDA_CHECK(expr)                                                                            \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      ::ppasr::set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                              std::to_string(__LINE__) + ")");                                           \
      return PPASR_ERR_CUDA;                                                                              \
    }                                                                                                     \
  } while (0)

#define PPASR_REQUIRE(cond, msg)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      ::ppasr::set_last_error(std::string("invalid argument: ") + (msg)); \
      return PPASR_ERR_INVALID;                                    \
    }                                                              \
  } while (0)
