// Fused per-channel bias + residual + activation for channels-last bf16 tensors (memory-bound, one pass).
//
// Used around the cuDNN 2-D convolutions that stay in PyTorch (encoders, ANM normal convs): with BatchNorm folded into
// the convolution, "conv -> +bias -> (+skip) -> PReLU/ReLU/LeakyReLU" is one read and one write instead of the three
// elementwise passes PyTorch launches (aten::add_, aten::add, aten::prelu).  The output may be a channel window of a
// wider tensor (y_cstride / y_coff), which also replaces torch.cat of the DPBlock dilated branches
// (src/model/stereodpnet/modules.py:42-44 of the reference).
//   y[p, y_coff + c] = act( x[p, c] + bias[c] + res[p, c] ),  act(v) = v > 0 ? v : slope * v   (slope 0 = ReLU, 1 = none)
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include <algorithm>
#include "dpf_ptx.cuh"

namespace {

using namespace dpf;

__global__ void __launch_bounds__(256) bias_act_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ bias,
                                                       const __nv_bfloat16* __restrict__ res, __nv_bfloat16* __restrict__ y,
                                                       long long npix, int C, int y_cstride, int y_coff, float slope) {
  const int c8n = C >> 3;
  const long long total = npix * c8n;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(q % c8n);
    const long long pix = q / c8n;
    const uint4 u = ld_nc_v4(x + pix * C + c8 * 8);
    float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
    if (bias != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8 + 4));
      f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
    }
    if (res != nullptr) {
      const uint4 r = ld_nc_v4(res + pix * C + c8 * 8);
      f[0] += bf16_lo(r.x); f[1] += bf16_hi(r.x); f[2] += bf16_lo(r.y); f[3] += bf16_hi(r.y);
      f[4] += bf16_lo(r.z); f[5] += bf16_hi(r.z); f[6] += bf16_lo(r.w); f[7] += bf16_hi(r.w);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = f[k] > 0.f ? f[k] : slope * f[k];
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(y + pix * y_cstride + y_coff + c8 * 8) = o;
  }
}

}  // namespace

extern "C" int dpf_bias_act(const void* x, const float* bias, const void* res, void* y, long long npix, int C, int y_cstride,
                            int y_coff, float slope, void* stream) {
  DPF_REQUIRE(x && y, "dpf_bias_act: null pointer");
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(y) && (res == nullptr || DPF_ALIGNED16(res)) && (bias == nullptr || DPF_ALIGNED16(bias)),
              "dpf_bias_act: pointers must be 16-byte aligned");
  DPF_REQUIRE(C >= 8 && C % 8 == 0 && y_cstride % 8 == 0 && y_coff % 8 == 0 && y_coff + C <= y_cstride && npix > 0,
              "dpf_bias_act: bad channel layout C=%d y_cstride=%d y_coff=%d", C, y_cstride, y_coff);
  const long long total = npix * (C / 8);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(dpf::sm_count()) * 16));
  bias_act_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), bias, reinterpret_cast<const __nv_bfloat16*>(res),
      reinterpret_cast<__nv_bfloat16*>(y), npix, C, y_cstride, y_coff, slope);
  return dpf::after_launch("dpf_bias_act");
}

// ------------------------------------------------------------------------------------------------------------
// ANM tail: bilinear x4 upsample (align_corners=True) -> sigmoid -> mean over the K sampled planes -> *2-1, fused.
// Replaces final_layer (upsampler + Sigmoid) and the mean / rescale of ANM.forward
// (src/model/stereodpnet/normal_module.py:69-72,185-190): reads [B*K,H4,W4,3] bf16 once, writes [B,3,H,W] fp32 once
// (the reference materialises the [B*K,3,H,W] tensor twice).
// ------------------------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) anm_tail_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int K,
                                                       int H4loc, int W4, int cs, int H4glob, int q_row0, int Hout, int y_row0) {
  const int H = 4 * H4glob, W = 4 * W4;                       // GLOBAL full-resolution size (align_corners scale)
  const int xo = blockIdx.x * 32 + (threadIdx.x & 31);
  const int yl = blockIdx.y * 8 + (threadIdx.x >> 5);        // local output row
  const int b = blockIdx.z;
  if (xo >= W || yl >= Hout) return;
  const int yo = y_row0 + yl;
  const float sh = static_cast<float>(H4glob - 1) / static_cast<float>(H - 1);
  const float sw = static_cast<float>(W4 - 1) / static_cast<float>(W - 1);
  const float sy = sh * static_cast<float>(yo), sx = sw * static_cast<float>(xo);
  const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
  const int y1 = y0 + (y0 < H4glob - 1 ? 1 : 0), x1 = x0 + (x0 < W4 - 1 ? 1 : 0);
  const float ly = sy - static_cast<float>(y0), lx = sx - static_cast<float>(x0);
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const int r0 = min(max(y0 - q_row0, 0), H4loc - 1), r1 = min(max(y1 - q_row0, 0), H4loc - 1);   // rows of the local tile
  float acc[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < K; ++k) {
    const __nv_bfloat16* p = x + static_cast<size_t>(b * K + k) * H4loc * W4 * cs;
    const __nv_bfloat16* p00 = p + (static_cast<size_t>(r0) * W4 + x0) * cs;
    const __nv_bfloat16* p01 = p + (static_cast<size_t>(r0) * W4 + x1) * cs;
    const __nv_bfloat16* p10 = p + (static_cast<size_t>(r1) * W4 + x0) * cs;
    const __nv_bfloat16* p11 = p + (static_cast<size_t>(r1) * W4 + x1) * cs;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = w00 * __bfloat162float(p00[c]) + w01 * __bfloat162float(p01[c]) + w10 * __bfloat162float(p10[c]) +
                      w11 * __bfloat162float(p11[c]);
      acc[c] += 1.0f / (1.0f + __expf(-v));
    }
  }
  const float inv = 2.0f / static_cast<float>(K);
#pragma unroll
  for (int c = 0; c < 3; ++c) out[((static_cast<size_t>(b) * 3 + c) * Hout + yl) * W + xo] = acc[c] * inv - 1.0f;
}

}  // namespace

extern "C" int dpf_anm_tail_tile(const void* x, float* out, int B, int K, int H4loc, int W4, int x_cstride, int H4glob, int q_row0,
                                 int Hout, int y_row0, void* stream) {
  DPF_REQUIRE(x && out, "dpf_anm_tail: null pointer");
  DPF_REQUIRE(B > 0 && B <= 65535 && K >= 1 && H4loc >= 1 && W4 > 1 && H4glob > 1 && x_cstride >= 3 && Hout >= 1, "dpf_anm_tail: bad shape");
  DPF_REQUIRE(y_row0 >= 0 && y_row0 + Hout <= 4 * H4glob, "dpf_anm_tail: output rows outside the image");
  {   // every quarter-res row the output rows interpolate from must be inside the local tile (halo rows may lie outside the image)
    const float sh = static_cast<float>(H4glob - 1) / static_cast<float>(4 * H4glob - 1);
    const int lo = static_cast<int>(sh * static_cast<float>(y_row0));
    const int hi = std::min(static_cast<int>(sh * static_cast<float>(y_row0 + Hout - 1)) + 1, H4glob - 1);
    DPF_REQUIRE(lo >= q_row0 && hi < q_row0 + H4loc, "dpf_anm_tail: the tile (rows %d..%d) does not cover rows %d..%d", q_row0,
                q_row0 + H4loc - 1, lo, hi);
  }
  dim3 grid((4 * W4 + 31) / 32, (Hout + 7) / 8, B);
  anm_tail_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), out, K, H4loc, W4,
                                                                      x_cstride, H4glob, q_row0, Hout, y_row0);
  return dpf::after_launch("dpf_anm_tail");
}

extern "C" int dpf_anm_tail(const void* x, float* out, int B, int K, int H4, int W4, void* stream) {
  return dpf_anm_tail_tile(x, out, B, K, H4, W4, 3, H4, 0, 4 * H4, 0, stream);
}

// ------------------------------------------------------------------------------------------------------------
// Feature-pyramid tail of the StereoDPNet encoder (src/model/stereodpnet/modules.py:128-133 of the reference):
//   cat([f1, bilinear_x2(f2), bilinear_x4(f3)], channel)  with align_corners=True, channels-last bf16.
// One pass writes the [N,h,w,3C] tensor directly (PyTorch: two upsample kernels, a cat and a layout copy).
// ------------------------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) pyramid_cat_kernel(const __nv_bfloat16* __restrict__ f1, const __nv_bfloat16* __restrict__ f2,
                                                          const __nv_bfloat16* __restrict__ f3, __nv_bfloat16* __restrict__ out,
                                                          int N, int h, int w, int h2, int w2, int h3, int w3, int C, int hg, int row0) {
  // one thread per (pixel, pyramid level): the source coordinates and blend weights are computed once and reused for all
  // C/8 channel chunks; the three threads of a pixel write its 3*C output channels contiguously.
  // Row crops (config 5): the maps hold rows row0 .. (level 1), row0/2 .. , row0/4 .. of images hg, hg/2, hg/4 rows tall; the
  // source rows follow the GLOBAL align_corners scale.  Untiled: hg = h, row0 = 0.
  const int c8n = C >> 3;
  const long long total = static_cast<long long>(N) * h * w * 3;
  const int hg2 = hg / 2, hg3 = hg / 4;
  const float ry2 = (hg > 1) ? static_cast<float>(hg2 - 1) / static_cast<float>(hg - 1) : 0.f;
  const float rx2 = (w > 1) ? static_cast<float>(w2 - 1) / static_cast<float>(w - 1) : 0.f;
  const float ry3 = (hg > 1) ? static_cast<float>(hg3 - 1) / static_cast<float>(hg - 1) : 0.f;
  const float rx3 = (w > 1) ? static_cast<float>(w3 - 1) / static_cast<float>(w - 1) : 0.f;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int lvl = static_cast<int>(q % 3);
    long long t = q / 3;
    const int x = static_cast<int>(t % w);
    t /= w;
    const int y = static_cast<int>(t % h);
    const int n = static_cast<int>(t / h);
    __nv_bfloat16* dst = out + ((static_cast<size_t>(n) * h + y) * w + x) * (3 * C) + lvl * C;
    if (lvl == 0) {
      const __nv_bfloat16* src = f1 + ((static_cast<size_t>(n) * h + y) * w + x) * C;
      for (int c = 0; c < c8n; ++c) *reinterpret_cast<uint4*>(dst + c * 8) = dpf::ld_nc_v4(src + c * 8);
      continue;
    }
    const __nv_bfloat16* src = lvl == 1 ? f2 : f3;
    const int hs = lvl == 1 ? h2 : h3, ws = lvl == 1 ? w2 : w3;
    const float sy = (lvl == 1 ? ry2 : ry3) * static_cast<float>(y + row0), sx = (lvl == 1 ? rx2 : rx3) * static_cast<float>(x);
    const int hgs = lvl == 1 ? hg2 : hg3, rs = lvl == 1 ? row0 / 2 : row0 / 4;         // global height / first row of the source level
    const int gy0 = min(static_cast<int>(sy), hgs - 1), gy1 = min(gy0 + 1, hgs - 1);   // global source rows
    const float ly = sy - static_cast<float>(gy0);
    const int y0 = min(max(gy0 - rs, 0), hs - 1), y1 = min(max(gy1 - rs, 0), hs - 1);  // rows of the local crop
    const int x0 = min(static_cast<int>(sx), ws - 1), x1 = min(x0 + 1, ws - 1);
    const float lx = sx - static_cast<float>(x0);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const __nv_bfloat16* base = src + static_cast<size_t>(n) * hs * ws * C;
    const __nv_bfloat16* p00 = base + (static_cast<size_t>(y0) * ws + x0) * C;
    const __nv_bfloat16* p01 = base + (static_cast<size_t>(y0) * ws + x1) * C;
    const __nv_bfloat16* p10 = base + (static_cast<size_t>(y1) * ws + x0) * C;
    const __nv_bfloat16* p11 = base + (static_cast<size_t>(y1) * ws + x1) * C;
    for (int c = 0; c < c8n; ++c) {
      const uint4 a = dpf::ld_nc_v4(p00 + c * 8), b = dpf::ld_nc_v4(p01 + c * 8), cc = dpf::ld_nc_v4(p10 + c * 8), d = dpf::ld_nc_v4(p11 + c * 8);
      const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w}, uc[4] = {cc.x, cc.y, cc.z, cc.w}, ud[4] = {d.x, d.y, d.z, d.w};
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o[k] = dpf::pack_bf16x2(w00 * dpf::bf16_lo(ua[k]) + w01 * dpf::bf16_lo(ub[k]) + w10 * dpf::bf16_lo(uc[k]) + w11 * dpf::bf16_lo(ud[k]),
                                w00 * dpf::bf16_hi(ua[k]) + w01 * dpf::bf16_hi(ub[k]) + w10 * dpf::bf16_hi(uc[k]) + w11 * dpf::bf16_hi(ud[k]));
      *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// per-pixel maximum over the channels of a channels-last bf16 map (ref_feature = ref_fea.max(1)[0], mainmodel.py:104)
__global__ void __launch_bounds__(256) channel_max_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long npix, int C) {
  const int c8n = C >> 3;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < npix;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    float m = -INFINITY;
    for (int c = 0; c < c8n; ++c) {
      const uint4 u = dpf::ld_nc_v4(x + q * C + c * 8);
      m = fmaxf(m, fmaxf(fmaxf(fmaxf(dpf::bf16_lo(u.x), dpf::bf16_hi(u.x)), fmaxf(dpf::bf16_lo(u.y), dpf::bf16_hi(u.y))),
                         fmaxf(fmaxf(dpf::bf16_lo(u.z), dpf::bf16_hi(u.z)), fmaxf(dpf::bf16_lo(u.w), dpf::bf16_hi(u.w)))));
    }
    y[q] = m;
  }
}

}  // namespace

extern "C" int dpf_pyramid_cat_tile(const void* f1, const void* f2, const void* f3, void* out, int N, int h, int w, int h2, int w2,
                                    int h3, int w3, int C, int hglob, int row0, void* stream);

extern "C" int dpf_pyramid_cat(const void* f1, const void* f2, const void* f3, void* out, int N, int h, int w, int h2, int w2,
                               int h3, int w3, int C, void* stream) {
  DPF_REQUIRE(h2 * 2 == h && h3 * 4 == h, "dpf_pyramid_cat: levels must be h, h/2, h/4 rows");
  return dpf_pyramid_cat_tile(f1, f2, f3, out, N, h, w, h2, w2, h3, w3, C, h, 0, stream);
}

extern "C" int dpf_pyramid_cat_tile(const void* f1, const void* f2, const void* f3, void* out, int N, int h, int w, int h2, int w2,
                                    int h3, int w3, int C, int hglob, int row0, void* stream) {
  DPF_REQUIRE(f1 && f2 && f3 && out, "dpf_pyramid_cat: null pointer");
  DPF_REQUIRE(hglob % 4 == 0 && row0 % 4 == 0 && row0 >= 0 && row0 + h <= hglob && h2 * 2 == h && h3 * 4 == h,
              "dpf_pyramid_cat: a crop must start on a multiple of 4 rows and the levels must be h, h/2, h/4 rows");
  DPF_REQUIRE(DPF_ALIGNED16(f1) && DPF_ALIGNED16(f2) && DPF_ALIGNED16(f3) && DPF_ALIGNED16(out), "dpf_pyramid_cat: pointers must be 16-byte aligned");
  DPF_REQUIRE(N > 0 && h > 0 && w > 0 && h2 > 0 && w2 > 0 && h3 > 0 && w3 > 0 && C >= 8 && C % 8 == 0, "dpf_pyramid_cat: bad shape");
  const long long total = static_cast<long long>(N) * h * w * 3;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(dpf::sm_count()) * 16));
  pyramid_cat_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(f1), reinterpret_cast<const __nv_bfloat16*>(f2), reinterpret_cast<const __nv_bfloat16*>(f3),
      reinterpret_cast<__nv_bfloat16*>(out), N, h, w, h2, w2, h3, w3, C, hglob, row0);
  return dpf::after_launch("dpf_pyramid_cat");
}

// ------------------------------------------------------------------------------------------------------------
// FPN top-down merge (torchvision.ops.FeaturePyramidNetwork.forward as used by src/model/stereodpnet/modules.py:83,124):
//   y[n, p, c] = x[n, p, c] + bias[c] + top[n, nearest(p), c]      (lateral 1x1 conv output + bias + nearest-upsampled coarser level)
// one pass instead of aten::add_ (bias), upsample_nearest2d and aten::add.
// ------------------------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) fpn_merge_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ bias,
                                                        const __nv_bfloat16* __restrict__ top, __nv_bfloat16* __restrict__ y, int N,
                                                        int h, int w, int ht, int wt, int C) {
  const int c8n = C >> 3;
  const long long total = static_cast<long long>(N) * h * w * c8n;
  const float sy = static_cast<float>(ht) / static_cast<float>(h), sx = static_cast<float>(wt) / static_cast<float>(w);
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(q % c8n) * 8;
    long long t = q / c8n;
    const int px = static_cast<int>(t % w);
    t /= w;
    const int py = static_cast<int>(t % h);
    const int n = static_cast<int>(t / h);
    const int ty = min(static_cast<int>(floorf(static_cast<float>(py) * sy)), ht - 1);
    const int tx = min(static_cast<int>(floorf(static_cast<float>(px) * sx)), wt - 1);
    const size_t o = ((static_cast<size_t>(n) * h + py) * w + px) * C + c0;
    const uint4 u = dpf::ld_nc_v4(x + o);
    const uint4 r = dpf::ld_nc_v4(top + ((static_cast<size_t>(n) * ht + ty) * wt + tx) * C + c0);
    float f[8] = {dpf::bf16_lo(u.x), dpf::bf16_hi(u.x), dpf::bf16_lo(u.y), dpf::bf16_hi(u.y),
                  dpf::bf16_lo(u.z), dpf::bf16_hi(u.z), dpf::bf16_lo(u.w), dpf::bf16_hi(u.w)};
    if (bias != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
      f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
    }
    // the lateral (x + bias) is rounded to bf16 first, like the conv output of the PyTorch module it replaces
    uint4 l;
    l.x = dpf::pack_bf16x2(f[0], f[1]); l.y = dpf::pack_bf16x2(f[2], f[3]); l.z = dpf::pack_bf16x2(f[4], f[5]); l.w = dpf::pack_bf16x2(f[6], f[7]);
    uint4 out;
    out.x = dpf::pack_bf16x2(dpf::bf16_lo(l.x) + dpf::bf16_lo(r.x), dpf::bf16_hi(l.x) + dpf::bf16_hi(r.x));
    out.y = dpf::pack_bf16x2(dpf::bf16_lo(l.y) + dpf::bf16_lo(r.y), dpf::bf16_hi(l.y) + dpf::bf16_hi(r.y));
    out.z = dpf::pack_bf16x2(dpf::bf16_lo(l.z) + dpf::bf16_lo(r.z), dpf::bf16_hi(l.z) + dpf::bf16_hi(r.z));
    out.w = dpf::pack_bf16x2(dpf::bf16_lo(l.w) + dpf::bf16_lo(r.w), dpf::bf16_hi(l.w) + dpf::bf16_hi(r.w));
    *reinterpret_cast<uint4*>(y + o) = out;
  }
}

}  // namespace

extern "C" int dpf_channel_max(const void* x, float* y, long long npix, int C, void* stream) {
  DPF_REQUIRE(x && y && npix > 0 && C >= 8 && C % 8 == 0 && DPF_ALIGNED16(x), "dpf_channel_max: bad arguments (C=%d)", C);
  const int blocks = static_cast<int>(std::min<long long>((npix + 255) / 256, static_cast<long long>(dpf::sm_count()) * 16));
  channel_max_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), y, npix, C);
  return dpf::after_launch("dpf_channel_max");
}

extern "C" int dpf_fpn_merge(const void* x, const float* bias, const void* top, void* y, int N, int h, int w, int ht, int wt, int C,
                             void* stream) {
  DPF_REQUIRE(x && top && y, "dpf_fpn_merge: null pointer");
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(top) && DPF_ALIGNED16(y) && (bias == nullptr || DPF_ALIGNED16(bias)),
              "dpf_fpn_merge: pointers must be 16-byte aligned");
  DPF_REQUIRE(N > 0 && h > 0 && w > 0 && ht > 0 && wt > 0 && C >= 8 && C % 8 == 0, "dpf_fpn_merge: bad shape");
  const long long total = static_cast<long long>(N) * h * w * (C / 8);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(dpf::sm_count()) * 16));
  fpn_merge_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), bias, reinterpret_cast<const __nv_bfloat16*>(top), reinterpret_cast<__nv_bfloat16*>(y),
      N, h, w, ht, wt, C);
  return dpf::after_launch("dpf_fpn_merge");
}
