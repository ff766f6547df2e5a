// Shared host-side helpers of libdpf_sm100.so: last-error string, launch counter, argument checks.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace dpf {

inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return 1;
}
inline std::atomic<long long>& launch_counter() {
  static std::atomic<long long> c{0};
  return c;
}
inline int after_launch(const char* what) {
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("%s: launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}
inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
#define DPF_REQUIRE(cond, ...) \
  do {                         \
    if (!(cond)) return ::dpf::fail(__VA_ARGS__); \
  } while (0)
#define DPF_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)

}  // namespace dpf
