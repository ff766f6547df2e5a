// Backward of the ASM (adaptive sampling) volume pieces, sm_100a (memory-bound).
//
//  dpf_asm_blend_bwd  : gradient of  y = mean_s( x_s * softmax_s( sigmoid( l_s*a + d ) ) )   (src/module/asm/asm.py:160-171)
//                       wrt the samples x_s (direct path) and wrt the normalised logits lhat_s = l_s*a + d.
//                       dy is summed over the D_rep volume slices that the forward wrote (cached-first-level mode).
//  dpf_asm_sample_bwd : transpose of the table-driven gather of dpf_asm_sample_fwd (src/module/asm/asm.py:87-127):
//                       dfeat[b, ri, ci, :] += rw*cw * dsamples[b,s,h,w,:]   (fp32 accumulation buffer, atomics;
//                       the feature map is 8x smaller than the volume, contention is low).
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

template <int S>
__global__ void __launch_bounds__(256) asm_blend_bwd_kernel(const __nv_bfloat16* __restrict__ samples,
                                                            const __nv_bfloat16* __restrict__ logits,
                                                            const float* __restrict__ in_a, const float* __restrict__ in_d,
                                                            const __nv_bfloat16* __restrict__ dvol,
                                                            __nv_bfloat16* __restrict__ dsamples, __nv_bfloat16* __restrict__ dlhat,
                                                            int B, int H, int W, int C, int Dvol, int d0, int D_rep, int ch_off,
                                                            int Cvol) {
  const int c8n = C >> 3;
  const long long total = static_cast<long long>(B) * H * W * c8n;
  const size_t plane = static_cast<size_t>(H) * W;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(q % c8n);
    const long long pix = q / c8n;
    const int b = static_cast<int>(pix / plane);
    const size_t hw = static_cast<size_t>(pix % plane);
    float dy[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int d = d0; d < d0 + D_rep; ++d) {
      float t[8];
      unpack8(ld_nc_v4(dvol + ((static_cast<size_t>(b) * Dvol + d) * plane + hw) * Cvol + ch_off + c8 * 8), t);
#pragma unroll
      for (int k = 0; k < 8; ++k) dy[k] += t[k];
    }
    float xs[S][8], gs[S][8], ps[S][8];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const size_t off = ((static_cast<size_t>(b) * S + s) * plane + hw) * C + c8 * 8;
      unpack8(ld_nc_v4(samples + off), xs[s]);
      float l[8];
      unpack8(ld_nc_v4(logits + off), l);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        gs[s][k] = 1.0f / (1.0f + __expf(-(l[k] * in_a[b * C + c8 * 8 + k] + in_d[b * C + c8 * 8 + k])));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float m = gs[0][k];
#pragma unroll
      for (int s = 1; s < S; ++s) m = fmaxf(m, gs[s][k]);
      float den = 0.f;
#pragma unroll
      for (int s = 0; s < S; ++s) { ps[s][k] = __expf(gs[s][k] - m); den += ps[s][k]; }
      const float inv = 1.0f / den;
      float dot = 0.f;
#pragma unroll
      for (int s = 0; s < S; ++s) { ps[s][k] *= inv; dot += ps[s][k] * xs[s][k]; }     // sum_t p_t * x_t
      const float g = dy[k] * (1.0f / static_cast<float>(S));
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float dp_minus = g * (xs[s][k] - dot);                      // dp_s - sum_t p_t dp_t, with dp_s = g * x_s
        const float dg = ps[s][k] * dp_minus;                             // softmax backward
        gs[s][k] = dg * gs[s][k] * (1.0f - gs[s][k]);                     // sigmoid backward -> d lhat_s  (reuse gs)
        xs[s][k] = g * ps[s][k];                                          // direct path -> d x_s           (reuse xs)
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const size_t off = ((static_cast<size_t>(b) * S + s) * plane + hw) * C + c8 * 8;
      *reinterpret_cast<uint4*>(dsamples + off) = pack8(xs[s]);
      *reinterpret_cast<uint4*>(dlhat + off) = pack8(gs[s]);
    }
  }
}

__global__ void __launch_bounds__(256) asm_sample_bwd_kernel(const __nv_bfloat16* __restrict__ dsamples, float* __restrict__ dfeat,
                                                             int B, int H, int W, int C, int S, const int* __restrict__ ri,
                                                             const float* __restrict__ rw, const int* __restrict__ ci,
                                                             const float* __restrict__ cw) {
  const int c8n = C >> 3;
  const long long total = static_cast<long long>(B) * S * H * W * c8n;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(q % c8n);
    long long t = q / c8n;
    const int w = static_cast<int>(t % W); t /= W;
    const int h = static_cast<int>(t % H); t /= H;
    const int s = static_cast<int>(t % S);
    const int b = static_cast<int>(t / S);
    float g[8];
    unpack8(ld_nc_v4(dsamples + q * 8), g);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = ri[(s * H + h) * 2 + i];
      const float wr = rw[(s * H + h) * 2 + i];
      if (r < 0 || wr == 0.f) continue;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = ci[(s * W + w) * 2 + j];
        const float wc = cw[(s * W + w) * 2 + j];
        if (c < 0 || wc == 0.f) continue;
        float* dst = dfeat + ((static_cast<size_t>(b) * H + r) * W + c) * C + c8 * 8;
        const float ww = wr * wc;
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(dst + k, ww * g[k]);
      }
    }
  }
}

inline int nblocks(long long total) {
  return static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(dpf::sm_count()) * 16));
}

}  // namespace

extern "C" int dpf_asm_blend_bwd(const void* samples, const void* logits, const float* in_a, const float* in_d, const void* dvol,
                                 void* dsamples, void* dlhat, int B, int H4, int W4, int C, int S, int D_vol, int d0, int D_rep,
                                 int ch_off, int Cvol, void* stream) {
  DPF_REQUIRE(samples && logits && in_a && in_d && dvol && dsamples && dlhat, "dpf_asm_blend_bwd: null pointer");
  DPF_REQUIRE(C >= 8 && C % 8 == 0 && ch_off % 8 == 0 && Cvol % 8 == 0 && ch_off + C <= Cvol, "dpf_asm_blend_bwd: bad channel layout");
  DPF_REQUIRE(S >= 1 && S <= 3, "dpf_asm_blend_bwd: S=%d must be 1..3", S);
  DPF_REQUIRE(D_rep >= 1 && d0 >= 0 && d0 + D_rep <= D_vol, "dpf_asm_blend_bwd: bad level range");
  const int blocks = nblocks(static_cast<long long>(B) * H4 * W4 * (C / 8));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto sp = reinterpret_cast<const __nv_bfloat16*>(samples);
  auto lg = reinterpret_cast<const __nv_bfloat16*>(logits);
  auto dv = reinterpret_cast<const __nv_bfloat16*>(dvol);
  auto ds = reinterpret_cast<__nv_bfloat16*>(dsamples);
  auto dl = reinterpret_cast<__nv_bfloat16*>(dlhat);
  if (S == 3) asm_blend_bwd_kernel<3><<<blocks, 256, 0, st>>>(sp, lg, in_a, in_d, dv, ds, dl, B, H4, W4, C, D_vol, d0, D_rep, ch_off, Cvol);
  else if (S == 2) asm_blend_bwd_kernel<2><<<blocks, 256, 0, st>>>(sp, lg, in_a, in_d, dv, ds, dl, B, H4, W4, C, D_vol, d0, D_rep, ch_off, Cvol);
  else asm_blend_bwd_kernel<1><<<blocks, 256, 0, st>>>(sp, lg, in_a, in_d, dv, ds, dl, B, H4, W4, C, D_vol, d0, D_rep, ch_off, Cvol);
  return dpf::after_launch("dpf_asm_blend_bwd");
}

extern "C" int dpf_asm_sample_bwd(const void* dsamples, float* dfeat, int B, int H4, int W4, int C, int S, const int* ri,
                                  const float* rw, const int* ci, const float* cw, void* stream) {
  DPF_REQUIRE(dsamples && dfeat && ri && rw && ci && cw, "dpf_asm_sample_bwd: null pointer");
  DPF_REQUIRE(C >= 8 && C % 8 == 0 && S >= 1 && S <= 8 && B > 0 && H4 > 0 && W4 > 0, "dpf_asm_sample_bwd: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(dfeat, 0, static_cast<size_t>(B) * H4 * W4 * C * sizeof(float), st);
  if (e != cudaSuccess) return dpf::fail("dpf_asm_sample_bwd: memset: %s", cudaGetErrorString(e));
  asm_sample_bwd_kernel<<<nblocks(static_cast<long long>(B) * S * H4 * W4 * (C / 8)), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(dsamples), dfeat, B, H4, W4, C, S, ri, rw, ci, cw);
  return dpf::after_launch("dpf_asm_sample_bwd");
}
