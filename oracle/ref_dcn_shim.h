// Force-included (-include) when compiling the UNMODIFIED reference sources of the 3-D deformable convolution
// (/root/reference/src/module/dcn3d/src/**) into oracle/_ref/DCN.so -- test infrastructure only.
//
// The reference calls AT_DISPATCH_FLOATING_TYPES(input.type(), ...) (src/cuda/deform_conv_cuda.cu:96,233); current PyTorch
// only accepts an at::ScalarType there.  Re-declare the macro so that it takes either, without touching the reference files.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>

namespace dpf_ref_shim {
inline at::ScalarType scalar_type_of(at::ScalarType t) { return t; }
inline at::ScalarType scalar_type_of(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace dpf_ref_shim

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(dpf_ref_shim::scalar_type_of(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
