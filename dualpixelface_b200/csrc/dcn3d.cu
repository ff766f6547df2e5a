// 3-D deformable convolution (D3D) for sm_100a -- see include/dpf_sm100.h (6).  Placeholder until the gather-producer
// tcgen05 kernel lands: fails loudly (no fallback).
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"

extern "C" int dpf_dcn3d_fwd(const void* x, const float* offset, const void* w, const float* scale, const float* shift,
                             void* y, int B, int D, int H, int W, int Cin_pad, int Cout, int relu, void* stream) {
  (void)x; (void)offset; (void)w; (void)scale; (void)shift; (void)y; (void)B; (void)D; (void)H; (void)W; (void)Cin_pad;
  (void)Cout; (void)relu; (void)stream;
  return dpf::fail("dpf_dcn3d_fwd: not built yet");
}
