// ABI housekeeping entry points of libdpf_sm100.so (see include/dpf_sm100.h).
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"

extern "C" {

int dpf_abi_version(void) { return DPF_ABI_VERSION; }

const char* dpf_last_error(void) { return dpf::err_buf(); }

long long dpf_launch_count(void) { return dpf::launch_counter().load(); }

int dpf_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return dpf::fail("no CUDA device: %s", cudaGetErrorString(e));
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return dpf::fail("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (p.major != 10) return dpf::fail("libdpf_sm100 needs an sm_100a device (B200); found sm_%d%d (%s)", p.major, p.minor, p.name);
  if (p.sharedMemPerBlockOptin < 200 * 1024) return dpf::fail("device offers only %zu B opt-in shared memory", p.sharedMemPerBlockOptin);
  return 0;
}

}  // extern "C"
