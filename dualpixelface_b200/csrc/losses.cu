// Fused training losses of the path: masked smooth-L1 over the n disparity heads + the element-wise cosine
// normal loss, forward value AND gradient in ONE pass over the full-resolution maps.
//
// Replaces SMOOTHL1Loss.forward ('given' conversion, disparity target; src/loss/depth/smoothL1.py:15-49) and COSINELoss.forward
// (masked branch, one prediction; src/loss/normal/cosine.py:35-55) of the reference, whose boolean-mask gathers `x[mask]` and
// per-head F.smooth_l1_loss calls are ~25 elementwise / index kernels (and as many again in their backward) over 3 + 6
// full-resolution maps:
//   smoothL1 = sum_i w_i * mean_{mask}( sl1(pred_i - gt) ),           sl1(d) = 0.5 d^2 if |d| < 1 else |d| - 0.5
//   cosine   = mean_{mask, c}( 1 - clamp( a_c * g_c / max(|a| |g|, 1e-6), -1, 1 ) ),  a = pred / max(|pred|, 1e-6), g likewise
//              (the similarity is element-wise over the 3 components -- the reference's quirk, cosine.py:18-26 -- not their sum)
// One thread per pixel reads mask, target and predictions once, writes the UNNORMALISED gradients (d sl1 / d pred_i, and
// d sum_c(1 - sim_c) / d pred_normal, both times the mask) and accumulates the partial sums; the block combines them in a
// fixed order and writes one partial row, a second tiny kernel adds the rows in order: deterministic, no floating-point atomics.
// The normalisation by the mask count and the loss weights are scalars applied by the caller (losses.FusedLossFn).
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include <algorithm>

namespace {

constexpr int kMaxHeads = 4;
constexpr int kThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums layout: [0..n) sum mask*sl1 per head | [n] sum mask (pixel count) | [n+1] sum mask * sum_c (1 - sim_c)
__global__ void __launch_bounds__(kThreads) fused_losses_kernel(const float* __restrict__ pred_depth, int n,
                                                                const float* __restrict__ gt_disp, const float* __restrict__ mask,
                                                                const float* __restrict__ pred_normal,
                                                                const float* __restrict__ gt_normal, float* __restrict__ g_depth,
                                                                float* __restrict__ g_normal, float* __restrict__ partial,
                                                                long long hw, long long npix) {
  float acc[kMaxHeads + 2];
#pragma unroll
  for (int i = 0; i < kMaxHeads + 2; ++i) acc[i] = 0.f;
  for (long long q = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; q < npix; q += static_cast<long long>(gridDim.x) * kThreads) {
    const long long b = q / hw, pix = q - b * hw;
    const float m = (mask == nullptr || mask[q] > 0.f) ? 1.f : 0.f;
    const float gt = gt_disp[q];
#pragma unroll
    for (int i = 0; i < kMaxHeads; ++i) {
      if (i < n) {
        const long long o = (b * n + i) * hw + pix;
        const float d = pred_depth[o] - gt;
        const float ad = fabsf(d);
        acc[i] += m * (ad < 1.f ? 0.5f * d * d : ad - 0.5f);
        g_depth[o] = m * (ad < 1.f ? d : (d > 0.f ? 1.f : -1.f));
      }
    }
    acc[kMaxHeads] += m;
    if (pred_normal != nullptr) {
      const long long o = b * 3 * hw + pix;
      const float p0 = pred_normal[o], p1 = pred_normal[o + hw], p2 = pred_normal[o + 2 * hw];
      float g0 = gt_normal[o], g1 = gt_normal[o + hw], g2 = gt_normal[o + 2 * hw];
      const float eps = 1e-6f;
      const float r = sqrtf(p0 * p0 + p1 * p1 + p2 * p2), rc = fmaxf(r, eps);
      const float a0 = p0 / rc, a1 = p1 / rc, a2 = p2 / rc;
      const float gc = fmaxf(sqrtf(g0 * g0 + g1 * g1 + g2 * g2), eps);
      g0 /= gc; g1 /= gc; g2 /= gc;
      const float na = sqrtf(a0 * a0 + a1 * a1 + a2 * a2), ng = sqrtf(g0 * g0 + g1 * g1 + g2 * g2);
      const float den0 = na * ng, den = fmaxf(den0, eps);
      const float u0 = a0 * g0 / den, u1 = a1 * g1 / den, u2 = a2 * g2 / den;
      acc[kMaxHeads + 1] += m * ((1.f - fminf(fmaxf(u0, -1.f), 1.f)) + (1.f - fminf(fmaxf(u1, -1.f), 1.f)) + (1.f - fminf(fmaxf(u2, -1.f), 1.f)));
      // gradient of L = sum_c (1 - clamp(u_c)) wrt pred (torch semantics: clamp passes the gradient on [-1, 1] inclusive)
      const float du0 = (u0 >= -1.f && u0 <= 1.f) ? -1.f : 0.f, du1 = (u1 >= -1.f && u1 <= 1.f) ? -1.f : 0.f,
                  du2 = (u2 >= -1.f && u2 <= 1.f) ? -1.f : 0.f;
      const float t = (den0 > eps && na > 0.f) ? (du0 * u0 + du1 * u1 + du2 * u2) / den * ng / na : 0.f;   // sum_c dL/du_c * u_c/den * dden/da_k = t * a_k
      const float da0 = du0 * g0 / den - t * a0, da1 = du1 * g1 / den - t * a1, da2 = du2 * g2 / den - t * a2;
      // a = pred / max(r, eps): da_k/dpred_j = delta_kj / rc - pred_k / rc^2 * (r > eps ? pred_j / r : 0)
      const float s = (r > eps) ? (da0 * p0 + da1 * p1 + da2 * p2) / (rc * rc * r) : 0.f;
      g_normal[o] = m * (da0 / rc - s * p0);
      g_normal[o + hw] = m * (da1 / rc - s * p1);
      g_normal[o + 2 * hw] = m * (da2 / rc - s * p2);
    }
  }
  // fixed-order block reduction: warp shuffle tree, then the 8 warp results in warp order
  __shared__ float red[kThreads / 32][kMaxHeads + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < kMaxHeads + 2; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kMaxHeads + 2) {
    float v = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) v += red[w][threadIdx.x];
    partial[blockIdx.x * (kMaxHeads + 2) + threadIdx.x] = v;
  }
}

__global__ void fused_losses_final_kernel(const float* __restrict__ partial, float* __restrict__ sums, int nblocks, int n) {
  const int i = threadIdx.x;                 // 0 .. kMaxHeads+1
  if (i >= kMaxHeads + 2) return;
  float v = 0.f;
  for (int b = 0; b < nblocks; ++b) v += partial[b * (kMaxHeads + 2) + i];
  if (i < n) sums[i] = v;
  else if (i == kMaxHeads) sums[n] = v;
  else if (i == kMaxHeads + 1) sums[n + 1] = v;
}

inline int loss_blocks(long long npix) {
  return static_cast<int>(std::min<long long>((npix + kThreads - 1) / kThreads, static_cast<long long>(dpf::sm_count()) * 8));
}

}  // namespace

extern "C" long long dpf_fused_losses_ws_floats(long long npix) { return static_cast<long long>(loss_blocks(npix)) * (kMaxHeads + 2); }

extern "C" int dpf_fused_losses(const float* pred_depth, int n_heads, const float* gt_disp, const float* mask, const float* pred_normal,
                                const float* gt_normal, float* g_depth, float* g_normal, float* ws, float* sums, int B, int H, int W,
                                void* stream) {
  DPF_REQUIRE(pred_depth && gt_disp && g_depth && ws && sums, "dpf_fused_losses: null pointer");
  DPF_REQUIRE(n_heads >= 1 && n_heads <= kMaxHeads, "dpf_fused_losses: n_heads=%d not in [1,%d]", n_heads, kMaxHeads);
  DPF_REQUIRE((pred_normal == nullptr) || (gt_normal && g_normal), "dpf_fused_losses: the normal loss needs gt_normal and g_normal");
  DPF_REQUIRE(B > 0 && H > 0 && W > 0, "dpf_fused_losses: bad shape");
  const long long hw = static_cast<long long>(H) * W, npix = hw * B;
  const int blocks = loss_blocks(npix);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  fused_losses_kernel<<<blocks, kThreads, 0, st>>>(pred_depth, n_heads, gt_disp, mask, pred_normal, gt_normal, g_depth, g_normal, ws, hw, npix);
  if (int rc = dpf::after_launch("dpf_fused_losses")) return rc;
  fused_losses_final_kernel<<<1, 32, 0, st>>>(ws, sums, blocks, n_heads);
  return dpf::after_launch("dpf_fused_losses");
}
