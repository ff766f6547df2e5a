// Training-mode BatchNorm3d pieces around the tensor-core convolution (memory-bound, channels-last bf16).
//
// The reference trains with nn.BatchNorm3d in batch-statistics mode after every 3-D convolution
// (convbn_3d, src/module/asm/basics.py:32-36; used throughout src/model/stereodpnet/modules.py:204-337).  With z the raw
// convolution output (bf16), per channel c over all N = B*D*H*W positions:
//   forward : y = act( z * a[c] + b[c] + res ),  a = gamma / sqrt(var + eps), b = beta - mean * a     (dpf_affine_act)
//             statistics from dpf_channel_stats (sum, sum of squares)
//   backward: g = dy * [y > 0]  (ReLU mask from the saved output; g = dy without ReLU)
//             S1 = sum g, S2 = sum g*z                                                            (dpf_bn_bwd_reduce)
//             dz = a * ( g - S1/N - (z - mean) * inv_std^2 * (S2 - mean*S1)/N ),  dres = g          (dpf_bn_bwd_apply)
//             dgamma = inv_std * (S2 - mean*S1), dbeta = S1   (host, C numbers)
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"

namespace {

using namespace dpf;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

__global__ void __launch_bounds__(256) affine_act_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                                                         const float* __restrict__ bias, const __nv_bfloat16* __restrict__ res,
                                                         __nv_bfloat16* __restrict__ y, long long npix, int C, float slope) {
  const int c8n = C >> 3;
  const long long total = npix * c8n;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(q % c8n);
    float f[8];
    unpack8(ld_nc_v4(x + q * 8), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = f[k] * (scale ? __ldg(scale + c8 * 8 + k) : 1.f) + (bias ? __ldg(bias + c8 * 8 + k) : 0.f);
    if (res != nullptr) {
      float r[8];
      unpack8(ld_nc_v4(res + q * 8), r);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += r[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = f[k] > 0.f ? f[k] : slope * f[k];
    *reinterpret_cast<uint4*>(y + q * 8) = pack8(f);
  }
}

// sums [2][C]: S1 = sum g, S2 = sum g*z.  grid (chunks), block 256 = (256/c8n) position lanes x c8n pieces.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                            const __nv_bfloat16* __restrict__ z, float* __restrict__ sums,
                                                            long long npix, int C, int relu, float slope) {
  extern __shared__ float red[];   // [2][C]
  const int c8n = C >> 3;
  const int piece = threadIdx.x % c8n, lanes = blockDim.x / c8n, pl = threadIdx.x / c8n;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float s1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, s2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (pl < lanes) {
    for (long long pos = static_cast<long long>(blockIdx.x) * lanes + pl; pos < npix; pos += static_cast<long long>(gridDim.x) * lanes) {
      const size_t off = static_cast<size_t>(pos) * C + piece * 8;
      float g[8], zz[8];
      unpack8(ld_nc_v4(dy + off), g);
      unpack8(ld_nc_v4(z + off), zz);
      if (relu) {
        float yy[8];
        unpack8(ld_nc_v4(y + off), yy);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = yy[k] > 0.f ? g[k] : slope * g[k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { s1[k] += g[k]; s2[k] = fmaf(g[k], zz[k], s2[k]); }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&red[piece * 8 + k], s1[k]);
      atomicAdd(&red[C + piece * 8 + k], s2[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&sums[i], red[i]);
}

// dz = A * (g - K1 - (z - MU) * K2);  dres = g (optional)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                           const __nv_bfloat16* __restrict__ z, const float* __restrict__ coef,
                                                           __nv_bfloat16* __restrict__ dz, __nv_bfloat16* __restrict__ dres,
                                                           long long npix, int C, int relu, float slope) {
  const int c8n = C >> 3;
  const long long total = npix * c8n;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(q % c8n);
    float g[8], zz[8], o[8];
    unpack8(ld_nc_v4(dy + q * 8), g);
    unpack8(ld_nc_v4(z + q * 8), zz);
    if (relu) {
      float yy[8];
      unpack8(ld_nc_v4(y + q * 8), yy);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = yy[k] > 0.f ? g[k] : slope * g[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c8 * 8 + k;
      o[k] = __ldg(coef + c) * (g[k] - __ldg(coef + C + c) - (zz[k] - __ldg(coef + 3 * C + c)) * __ldg(coef + 2 * C + c));
    }
    *reinterpret_cast<uint4*>(dz + q * 8) = pack8(o);
    if (dres != nullptr) *reinterpret_cast<uint4*>(dres + q * 8) = pack8(g);
  }
}

// Per-channel coefficients of a train-mode BatchNorm in ONE launch (the ~14 tiny element-wise launches this replaces cost more
// host time than the two bandwidth kernels around them): stats [C][2] = (sum z, sum z^2) over n elements per channel ->
// out [4][C] = a = gamma * inv_std | b = beta - mean * a | mean | inv_std; running statistics updated in place (momentum, unbiased var).
__global__ void bn_fwd_coefs_kernel(const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    float n, float eps, float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                    float* __restrict__ out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = stats[2 * c] / n;
  const float var = fmaxf(stats[2 * c + 1] / n - mean * mean, 0.f);
  const float inv = rsqrtf(var + eps);
  const float a = (gamma ? gamma[c] : 1.f) * inv;
  out[c] = a;
  out[C + c] = (beta ? beta[c] : 0.f) - mean * a;
  out[2 * C + c] = mean;
  out[3 * C + c] = inv;
  if (running_mean != nullptr) {
    running_mean[c] = running_mean[c] * (1.f - momentum) + momentum * mean;
    running_var[c] = running_var[c] * (1.f - momentum) + momentum * var * (n / fmaxf(n - 1.f, 1.f));
  }
}

// Backward coefficients: sums [2][C] = (S1 = sum g, S2 = sum g z) -> dgamma = inv_std (S2 - mean S1), dbeta = S1,
// coef [4][C] = a | S1 / n | inv_std^2 (S2 - mean S1) / n | mean   (the argument of dpf_bn_bwd_apply).
__global__ void bn_bwd_coefs_kernel(const float* __restrict__ sums, const float* __restrict__ fwd /*[4][C] of the forward*/, float n,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s1 = sums[c], s2 = sums[C + c];
  const float a = fwd[c], mean = fwd[2 * C + c], inv = fwd[3 * C + c];
  const float centred = s2 - mean * s1;
  dgamma[c] = inv * centred;
  dbeta[c] = s1;
  coef[c] = a;
  coef[C + c] = s1 / n;
  coef[2 * C + c] = inv * inv * centred / n;
  coef[3 * C + c] = mean;
}

inline int nblocks(long long total) {
  return static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(dpf::sm_count()) * 16));
}

}  // namespace

extern "C" int dpf_affine_act(const void* x, const float* scale, const float* bias, const void* res, void* y, long long npix,
                              int C, float slope, void* stream) {
  DPF_REQUIRE(x && y, "dpf_affine_act: null pointer");
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(y) && (!res || DPF_ALIGNED16(res)), "dpf_affine_act: pointers must be 16-byte aligned");
  DPF_REQUIRE(C >= 8 && C % 8 == 0 && npix > 0, "dpf_affine_act: bad shape");
  affine_act_kernel<<<nblocks(npix * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), scale, bias, reinterpret_cast<const __nv_bfloat16*>(res),
      reinterpret_cast<__nv_bfloat16*>(y), npix, C, slope);
  return dpf::after_launch("dpf_affine_act");
}

extern "C" int dpf_bn_bwd_reduce(const void* dy, const void* y, const void* z, float* sums, long long npix, int C, int relu,
                                 float slope, void* stream) {
  DPF_REQUIRE(dy && z && sums && (!relu || y), "dpf_bn_bwd_reduce: null pointer");
  DPF_REQUIRE(C >= 8 && C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0 && npix > 0, "dpf_bn_bwd_reduce: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(sums, 0, 2 * C * sizeof(float), st);
  if (e != cudaSuccess) return dpf::fail("dpf_bn_bwd_reduce: memset: %s", cudaGetErrorString(e));
  const int lanes = 256 / (C / 8);
  const int chunks = static_cast<int>(std::min<long long>((npix + lanes * 8 - 1) / (lanes * 8), static_cast<long long>(dpf::sm_count()) * 8));
  bn_bwd_reduce_kernel<<<chunks, 256, 2 * C * sizeof(float), st>>>(reinterpret_cast<const __nv_bfloat16*>(dy),
                                                                   reinterpret_cast<const __nv_bfloat16*>(y),
                                                                   reinterpret_cast<const __nv_bfloat16*>(z), sums, npix, C, relu, slope);
  return dpf::after_launch("dpf_bn_bwd_reduce");
}

extern "C" int dpf_bn_bwd_apply(const void* dy, const void* y, const void* z, const float* coef, void* dz, void* dres,
                                long long npix, int C, int relu, float slope, void* stream) {
  DPF_REQUIRE(dy && z && coef && dz && (!relu || y), "dpf_bn_bwd_apply: null pointer");
  DPF_REQUIRE(C >= 8 && C % 8 == 0 && npix > 0, "dpf_bn_bwd_apply: bad shape");
  bn_bwd_apply_kernel<<<nblocks(npix * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(y), reinterpret_cast<const __nv_bfloat16*>(z),
      coef, reinterpret_cast<__nv_bfloat16*>(dz), reinterpret_cast<__nv_bfloat16*>(dres), npix, C, relu, slope);
  return dpf::after_launch("dpf_bn_bwd_apply");
}

extern "C" int dpf_bn_fwd_coefs(const float* stats, const float* gamma, const float* beta, float n, float eps, float momentum,
                                float* running_mean, float* running_var, float* out, int C, void* stream) {
  DPF_REQUIRE(stats && out && C > 0 && n > 0.f, "dpf_bn_fwd_coefs: bad arguments");
  DPF_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "dpf_bn_fwd_coefs: running_mean and running_var go together");
  bn_fwd_coefs_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(stats, gamma, beta, n, eps, momentum, running_mean,
                                                                                      running_var, out, C);
  return dpf::after_launch("dpf_bn_fwd_coefs");
}

extern "C" int dpf_bn_bwd_coefs(const float* sums, const float* fwd, float n, float* dgamma, float* dbeta, float* coef, int C,
                                void* stream) {
  DPF_REQUIRE(sums && fwd && dgamma && dbeta && coef && C > 0 && n > 0.f, "dpf_bn_bwd_coefs: bad arguments");
  bn_bwd_coefs_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(sums, fwd, n, dgamma, dbeta, coef, C);
  return dpf::after_launch("dpf_bn_bwd_coefs");
}
