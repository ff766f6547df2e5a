// Backward of the two memory-bound ANM ops (normal branch, src/model/stereodpnet/normal_module.py:140-194):
//   dpf_anm_tail_bwd    adjoint of  mean_k(sigmoid(bilinear x4 (align_corners))) * 2 - 1   (normal_module.py:186-192)
//   dpf_anm_gather_bwd  adjoint of  the level gather that builds the sampled feature volume (normal_module.py:155-171):
//                       the cost channels of the gradient go back to the selected levels, every other level gets zero;
//                       the coordinate channels carry no gradient (the sampled levels are constants).
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"

namespace {

__global__ void __launch_bounds__(256) anm_tail_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ dout,
                                                           float* __restrict__ dx, int K, int H4, int W4) {
  const int H = 4 * H4, W = 4 * W4;
  const int xo = blockIdx.x * 32 + (threadIdx.x & 31);
  const int yo = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (xo >= W || yo >= H) return;
  const float sh = static_cast<float>(H4 - 1) / static_cast<float>(H - 1);
  const float sw = static_cast<float>(W4 - 1) / static_cast<float>(W - 1);
  const float sy = sh * static_cast<float>(yo), sx = sw * static_cast<float>(xo);
  const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
  const int y1 = y0 + (y0 < H4 - 1 ? 1 : 0), x1 = x0 + (x0 < W4 - 1 ? 1 : 0);
  const float ly = sy - static_cast<float>(y0), lx = sx - static_cast<float>(x0);
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const float inv = 2.0f / static_cast<float>(K);
  float g[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) g[c] = dout[((static_cast<size_t>(b) * 3 + c) * H + yo) * W + xo] * inv;
  const size_t o00 = (static_cast<size_t>(y0) * W4 + x0) * 3, o01 = (static_cast<size_t>(y0) * W4 + x1) * 3;
  const size_t o10 = (static_cast<size_t>(y1) * W4 + x0) * 3, o11 = (static_cast<size_t>(y1) * W4 + x1) * 3;
  for (int k = 0; k < K; ++k) {
    const size_t base = static_cast<size_t>(b * K + k) * H4 * W4 * 3;
    const __nv_bfloat16* p = x + base;
    float* q = dx + base;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = w00 * __bfloat162float(p[o00 + c]) + w01 * __bfloat162float(p[o01 + c]) + w10 * __bfloat162float(p[o10 + c]) +
                      w11 * __bfloat162float(p[o11 + c]);
      const float s = 1.0f / (1.0f + __expf(-v));
      const float gv = g[c] * s * (1.0f - s);
      atomicAdd(q + o00 + c, gv * w00);
      atomicAdd(q + o01 + c, gv * w01);
      atomicAdd(q + o10 + c, gv * w10);
      atomicAdd(q + o11 + c, gv * w11);
    }
  }
}

__global__ void __launch_bounds__(256) anm_gather_bwd_kernel(const float* __restrict__ dfv_a, const __nv_bfloat16* __restrict__ dfv_b,
                                                             const int* __restrict__ idx, __nv_bfloat16* __restrict__ dout3, int B,
                                                             int D, int K, int H4, int W4, int C, int Cpad) {
  const int pcs = C >> 3;
  const size_t n = static_cast<size_t>(H4) * W4;
  const long long total = static_cast<long long>(B) * D * n * pcs;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pc = static_cast<int>(q % pcs);
    long long t = q / pcs;
    const size_t pix = static_cast<size_t>(t % n);
    t /= n;
    const int d = static_cast<int>(t % D);
    const int b = static_cast<int>(t / D);
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < K; ++k) {                       // top-k levels are distinct, but summing keeps this a true adjoint
      const size_t src = (static_cast<size_t>(b) * K + k) * n + pix;
      if (idx[src] != d) continue;
      if (dfv_a) {
        const float4 a0 = *reinterpret_cast<const float4*>(dfv_a + src * Cpad + pc * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(dfv_a + src * Cpad + pc * 8 + 4);
        f[0] += a0.x; f[1] += a0.y; f[2] += a0.z; f[3] += a0.w; f[4] += a1.x; f[5] += a1.y; f[6] += a1.z; f[7] += a1.w;
      }
      if (dfv_b) {
        const uint4 u = *reinterpret_cast<const uint4*>(dfv_b + src * Cpad + pc * 8);
        f[0] += dpf::bf16_lo(u.x); f[1] += dpf::bf16_hi(u.x); f[2] += dpf::bf16_lo(u.y); f[3] += dpf::bf16_hi(u.y);
        f[4] += dpf::bf16_lo(u.z); f[5] += dpf::bf16_hi(u.z); f[6] += dpf::bf16_lo(u.w); f[7] += dpf::bf16_hi(u.w);
      }
    }
    uint4 o;
    o.x = dpf::pack_bf16x2(f[0], f[1]); o.y = dpf::pack_bf16x2(f[2], f[3]);
    o.z = dpf::pack_bf16x2(f[4], f[5]); o.w = dpf::pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(dout3 + ((static_cast<size_t>(b) * D + d) * n + pix) * C + pc * 8) = o;
  }
}

}  // namespace

extern "C" int dpf_anm_tail_bwd(const void* x, const float* dout, float* dx, int B, int K, int H4, int W4, void* stream) {
  DPF_REQUIRE(x && dout && dx, "dpf_anm_tail_bwd: null pointer");
  DPF_REQUIRE(B > 0 && B <= 65535 && K >= 1 && H4 > 1 && W4 > 1, "dpf_anm_tail_bwd: bad shape");
  dim3 grid((4 * W4 + 31) / 32, (4 * H4 + 7) / 8, B);
  anm_tail_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), dout, dx, K, H4, W4);
  return dpf::after_launch("dpf_anm_tail_bwd");
}

extern "C" int dpf_anm_gather_bwd(const float* dfv_f32, const void* dfv_bf16, const int* idx, void* dout3, int B, int D, int K,
                                  int H4, int W4, int C, int Cpad, void* stream) {
  DPF_REQUIRE((dfv_f32 || dfv_bf16) && idx && dout3, "dpf_anm_gather_bwd: null pointer");
  DPF_REQUIRE(C % 8 == 0 && Cpad % 8 == 0 && Cpad >= C, "dpf_anm_gather_bwd: need C %% 8 == 0 and Cpad >= C (C=%d Cpad=%d)", C, Cpad);
  DPF_REQUIRE(DPF_ALIGNED16(dout3) && (!dfv_f32 || DPF_ALIGNED16(dfv_f32)) && (!dfv_bf16 || DPF_ALIGNED16(dfv_bf16)),
              "dpf_anm_gather_bwd: pointers must be 16-byte aligned");
  const long long total = static_cast<long long>(B) * D * H4 * W4 * (C / 8);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(dpf::sm_count()) * 16));
  anm_gather_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dfv_f32, reinterpret_cast<const __nv_bfloat16*>(dfv_bf16), idx, reinterpret_cast<__nv_bfloat16*>(dout3), B, D, K, H4, W4, C, Cpad);
  return dpf::after_launch("dpf_anm_gather_bwd");
}
