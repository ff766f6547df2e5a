// Fused x4 trilinear upsample (align_corners=True) + softmax over 4*D bins + soft-argmin, forward and backward.
//
// Replaces F.interpolate(cost, scale_factor=4, mode='trilinear', align_corners=True)
// (src/model/stereodpnet/modules.py:327-334) followed by disp_regression.forward (modules.py:352-362) without ever
// materialising the [B,4D,H,W] tensor (963 MB fp32 per head at 1120x1680, batch 4).
// Coordinates follow ATen's area_pixel_compute_source_index(align_corners=true): src = dst * (in-1)/(out-1) in fp32,
// i0 = (int)src, i1 = i0 + (i0 < in-1), lambda1 = src - i0.
// Memory-bound: reads D*4 B per quarter-res pixel, writes 4 B per full-res pixel (+ 4*D*4 B if `prob` is requested).
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include <algorithm>

namespace {

constexpr int kMaxD = 16;
constexpr int kTX = 32, kTY = 8;   // full-res tile of one CTA (forward and backward)

struct Axis {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Axis src_index(int dst, float scale, int in_size) {
  Axis a;
  const float src = scale * static_cast<float>(dst);
  a.i0 = static_cast<int>(src);
  a.i1 = a.i0 + ((a.i0 < in_size - 1) ? 1 : 0);
  a.l1 = src - static_cast<float>(a.i0);
  a.l0 = 1.0f - a.l1;
  return a;
}

// ATen's area_pixel_compute_source_index(align_corners=false) for a x4 scale_factor: src = max(0.25 * (dst + 0.5) - 0.5, 0)
__device__ __forceinline__ Axis src_index_half(int dst, int in_size) {
  Axis a;
  const float src = fmaxf(0.25f * (static_cast<float>(dst) + 0.5f) - 0.5f, 0.f);
  a.i0 = min(static_cast<int>(src), in_size - 1);
  a.i1 = a.i0 + ((a.i0 < in_size - 1) ? 1 : 0);
  a.l1 = src - static_cast<float>(a.i0);
  a.l0 = 1.0f - a.l1;
  return a;
}

template <int D>
__device__ __forceinline__ void plane_values(const float* __restrict__ cost, size_t plane, int W4, const Axis& ay,
                                             const Axis& ax, float (&c)[D]) {
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float* p = cost + d * plane;
    const float v00 = __ldg(p + static_cast<size_t>(ay.i0) * W4 + ax.i0);
    const float v01 = __ldg(p + static_cast<size_t>(ay.i0) * W4 + ax.i1);
    const float v10 = __ldg(p + static_cast<size_t>(ay.i1) * W4 + ax.i0);
    const float v11 = __ldg(p + static_cast<size_t>(ay.i1) * W4 + ax.i1);
    c[d] = ay.l0 * (ax.l0 * v00 + ax.l1 * v01) + ay.l1 * (ax.l0 * v10 + ax.l1 * v11);
  }
}

// Row tiles (BASELINE config 5): `cost` holds the quarter-resolution rows q_row0 .. q_row0+H4loc-1 of an image that is H4 rows
// tall, `disp` / `prob` the full-resolution rows y_row0 .. y_row0+Hout-1; the source coordinates use the GLOBAL align_corners
// scale, so the tiles reproduce the untiled result exactly.  Untiled: H4loc = H4, q_row0 = y_row0 = 0, Hout = 4*H4.
// HALF: half-pixel (align_corners=False) coordinates on all three axes -- NNet's up-sampling (src/model/nnet/mainmodel.py:149-151).
template <int D, bool HALF = false>
__global__ void __launch_bounds__(kTX* kTY) regress_fwd_kernel(const float* __restrict__ cost, float* __restrict__ disp,
                                                               float* __restrict__ prob, int H4, int W4, float mindisp,
                                                               float step, int H4loc, int q_row0, int Hout, int y_row0) {
  const int H = 4 * H4, W = 4 * W4;
  const int x = blockIdx.x * kTX + threadIdx.x;
  const int yl = blockIdx.y * kTY + threadIdx.y;             // local output row
  const int b = blockIdx.z;
  if (x >= W || yl >= Hout) return;
  const int y = y_row0 + yl;
  const float sh = static_cast<float>(H4 - 1) / static_cast<float>(H - 1);
  const float sw = static_cast<float>(W4 - 1) / static_cast<float>(W - 1);
  const float sd = static_cast<float>(D - 1) / static_cast<float>(4 * D - 1);
  Axis ay = HALF ? src_index_half(y, H4) : src_index(y, sh, H4);
  const Axis ax = HALF ? src_index_half(x, W4) : src_index(x, sw, W4);
  ay.i0 = min(max(ay.i0 - q_row0, 0), H4loc - 1);             // global quarter-res rows -> rows of the local tile
  ay.i1 = min(max(ay.i1 - q_row0, 0), H4loc - 1);
  const size_t plane = static_cast<size_t>(H4loc) * W4;
  float c[D];
  plane_values<D>(cost + static_cast<size_t>(b) * D * plane, plane, W4, ay, ax, c);
  // The 4*D bins are linear interpolants of the D knots, so their maximum is a knot: the softmax shift m needs D compares, not
  // 4*D (the softmax is shift-invariant; any m >= max keeps the exponentials in range).  Knots go to the log2 domain once
  // (c <- (c - m) * log2 e), each bin is then one FFMA + one EX2, and the expectation uses sum_k k * e_k with immediate k.
  float m = c[0];
#pragma unroll
  for (int d = 1; d < D; ++d) m = fmaxf(m, c[d]);
#pragma unroll
  for (int d = 0; d < D; ++d) c[d] = (c[d] - m) * 1.4426950408889634f;
  // bin k in the log2 domain (d0, d1, l1 fold to constants once the loops below are unrolled)
  auto bin = [&](int k) {
    const float src = HALF ? fmaxf(0.25f * (static_cast<float>(k) + 0.5f) - 0.5f, 0.f) : sd * static_cast<float>(k);
    const int d0 = min(static_cast<int>(src), D - 1);
    const int d1 = d0 + ((d0 < D - 1) ? 1 : 0);
    const float l1 = src - static_cast<float>(d0);
    return fmaf(l1, c[d1] - c[d0], c[d0]);
  };
  float v[4 * D];
  float se = 0.f, sek = 0.f;
#pragma unroll
  for (int k = 0; k < 4 * D; ++k) {
    const float e = exp2f(bin(k));
    v[k] = e;
    se += e;
    sek = fmaf(e, static_cast<float>(k), sek);
  }
  if (!(se >= 1e-30f)) {
    // Degenerate costs (adjacent levels > ~100 apart, e.g. an uncalibrated network): interior knots are never hit exactly by a bin,
    // so every bin can underflow against the maximum KNOT.  Redo the sums against the maximum BIN, as torch's softmax does.
    float mb = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4 * D; ++k) mb = fmaxf(mb, bin(k));
    se = 0.f; sek = 0.f;
#pragma unroll
    for (int k = 0; k < 4 * D; ++k) {
      const float e = exp2f(bin(k) - mb);
      v[k] = e;
      se += e;
      sek = fmaf(e, static_cast<float>(k), sek);
    }
  }
  const float sed = fmaf(step, sek, mindisp * se);
  const float inv = 1.0f / se;
  disp[(static_cast<size_t>(b) * Hout + yl) * W + x] = sed * inv;
  if (prob != nullptr) {
    float* pp = prob + (static_cast<size_t>(b) * 4 * D * Hout + yl) * W + x;
#pragma unroll
    for (int k = 0; k < 4 * D; ++k) pp[static_cast<size_t>(k) * Hout * W] = v[k] * inv;
  }
}

// Backward: recompute the softmax per full-res pixel, push d(disp)/d(v_k) back through the depth and bilinear
// interpolation into a shared-memory tile of quarter-res cells, then one global reduction per touched cell.
template <int D>
__global__ void __launch_bounds__(kTX* kTY) regress_bwd_kernel(const float* __restrict__ cost,
                                                               const float* __restrict__ ddisp, float* __restrict__ dcost,
                                                               int H4, int W4, float mindisp, float step) {
  constexpr int CX = kTX / 4 + 2, CY = kTY / 4 + 2;          // quarter-res cells a tile can touch
  // one accumulator copy per warp (= per pixel row of the tile): shared-memory atomics then only contend inside a warp (~4 lanes
  // per cell) instead of across all 256 threads of the block
  __shared__ float acc[kTY][D][CY][CX];
  const int H = 4 * H4, W = 4 * W4;
  const int tid = threadIdx.y * kTX + threadIdx.x;
  for (int i = tid; i < kTY * D * CY * CX; i += kTX * kTY) (&acc[0][0][0][0])[i] = 0.f;
  __syncthreads();
  const int x = blockIdx.x * kTX + threadIdx.x;
  const int y = blockIdx.y * kTY + threadIdx.y;
  const int b = blockIdx.z;
  const float sh = static_cast<float>(H4 - 1) / static_cast<float>(H - 1);
  const float sw = static_cast<float>(W4 - 1) / static_cast<float>(W - 1);
  const float sd = static_cast<float>(D - 1) / static_cast<float>(4 * D - 1);
  // origin of the tile's cell window (cell of the tile's first pixel)
  const int cx0 = static_cast<int>(sw * static_cast<float>(blockIdx.x * kTX));
  const int cy0 = static_cast<int>(sh * static_cast<float>(blockIdx.y * kTY));
  const size_t plane = static_cast<size_t>(H4) * W4;
  if (x < W && y < H) {
    const Axis ay = src_index(y, sh, H4), ax = src_index(x, sw, W4);
    float c[D];
    plane_values<D>(cost + static_cast<size_t>(b) * D * plane, plane, W4, ay, ax, c);
    float v[4 * D];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4 * D; ++k) {
      const float src = sd * static_cast<float>(k);
      const int d0 = static_cast<int>(src);
      const int d1 = d0 + ((d0 < D - 1) ? 1 : 0);
      const float l1 = src - static_cast<float>(d0);
      v[k] = (1.0f - l1) * c[d0] + l1 * c[d1];
      m = fmaxf(m, v[k]);
    }
    float se = 0.f, sed = 0.f;
#pragma unroll
    for (int k = 0; k < 4 * D; ++k) {
      v[k] = __expf(v[k] - m);
      se += v[k];
      sed = fmaf(v[k], mindisp + step * static_cast<float>(k), sed);
    }
    const float inv = 1.0f / se;
    const float dsp = sed * inv;
    const float g = ddisp[(static_cast<size_t>(b) * H + y) * W + x];
    float dc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) dc[d] = 0.f;
#pragma unroll
    for (int k = 0; k < 4 * D; ++k) {
      const float dv = g * v[k] * inv * (mindisp + step * static_cast<float>(k) - dsp);
      const float src = sd * static_cast<float>(k);
      const int d0 = static_cast<int>(src);
      const int d1 = d0 + ((d0 < D - 1) ? 1 : 0);
      const float l1 = src - static_cast<float>(d0);
      dc[d0] += (1.0f - l1) * dv;
      dc[d1] += l1 * dv;
    }
    const int ly0 = ay.i0 - cy0, ly1 = ay.i1 - cy0, lx0 = ax.i0 - cx0, lx1 = ax.i1 - cx0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      atomicAdd(&acc[threadIdx.y][d][ly0][lx0], dc[d] * ay.l0 * ax.l0);
      atomicAdd(&acc[threadIdx.y][d][ly0][lx1], dc[d] * ay.l0 * ax.l1);
      atomicAdd(&acc[threadIdx.y][d][ly1][lx0], dc[d] * ay.l1 * ax.l0);
      atomicAdd(&acc[threadIdx.y][d][ly1][lx1], dc[d] * ay.l1 * ax.l1);
    }
  }
  __syncthreads();
  for (int i = tid; i < D * CY * CX; i += kTX * kTY) {
    const int d = i / (CY * CX);
    const int r = (i / CX) % CY;
    const int cc = i % CX;
    const int gy = cy0 + r, gx = cx0 + cc;
    float val = 0.f;
#pragma unroll
    for (int wy = 0; wy < kTY; ++wy) val += acc[wy][d][r][cc];        // fixed order over the block's rows
    if (gy < H4 && gx < W4 && val != 0.f) atomicAdd(dcost + (static_cast<size_t>(b) * D + d) * plane + static_cast<size_t>(gy) * W4 + gx, val);
  }
}


// Soft-argmin over D levels at the resolution of the cost (no up-sampling): one thread per position, coalesced plane reads.
__global__ void __launch_bounds__(256) softargmin_kernel(const float* __restrict__ cost, float* __restrict__ disp, float* __restrict__ prob,
                                                         int D, long long P, float mindisp, float step) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= P) return;
  const float* c = cost + static_cast<long long>(blockIdx.y) * D * P + i;
  float m = -INFINITY;
  for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(c + d * P));
  float se = 0.f, sed = 0.f;
  for (int d = 0; d < D; ++d) {
    const float e = __expf(__ldg(c + d * P) - m);
    se += e;
    sed = fmaf(e, mindisp + step * static_cast<float>(d), sed);
  }
  const float inv = 1.0f / se;
  disp[static_cast<long long>(blockIdx.y) * P + i] = sed * inv;
  if (prob != nullptr) {
    float* pp = prob + static_cast<long long>(blockIdx.y) * D * P + i;
    for (int d = 0; d < D; ++d) pp[d * P] = __expf(__ldg(c + d * P) - m) * inv;
  }
}

}  // namespace

extern "C" int dpf_regress_fwd_tile(const float* cost, float* disp, float* prob, int B, int D, int H4loc, int W4, int H4glob,
                                    int q_row0, int Hout, int y_row0, float mindisp, float step, void* stream) {
  DPF_REQUIRE(cost && disp, "dpf_regress_fwd: null pointer");
  DPF_REQUIRE(B > 0 && B <= 65535 && H4loc >= 1 && H4glob > 1 && W4 > 1 && Hout >= 1, "dpf_regress_fwd: bad shape B=%d H4=%d W4=%d", B, H4glob, W4);
  DPF_REQUIRE(D == 8 || D == 4 || D == 16, "dpf_regress_fwd: D=%d not in {4,8,16}", D);
  DPF_REQUIRE(y_row0 >= 0 && y_row0 + Hout <= 4 * H4glob, "dpf_regress_fwd: output rows outside the image");
  {   // every quarter-res row the output rows interpolate from must be inside the local tile
    const float sh = static_cast<float>(H4glob - 1) / static_cast<float>(4 * H4glob - 1);
    const int lo = static_cast<int>(sh * static_cast<float>(y_row0));
    const int hi = std::min(static_cast<int>(sh * static_cast<float>(y_row0 + Hout - 1)) + 1, H4glob - 1);
    DPF_REQUIRE(lo >= q_row0 && hi < q_row0 + H4loc, "dpf_regress_fwd: the cost tile (rows %d..%d) does not cover rows %d..%d",
                q_row0, q_row0 + H4loc - 1, lo, hi);
  }
  dim3 grid((4 * W4 + kTX - 1) / kTX, (Hout + kTY - 1) / kTY, B), block(kTX, kTY);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (D == 8) regress_fwd_kernel<8><<<grid, block, 0, st>>>(cost, disp, prob, H4glob, W4, mindisp, step, H4loc, q_row0, Hout, y_row0);
  else if (D == 4) regress_fwd_kernel<4><<<grid, block, 0, st>>>(cost, disp, prob, H4glob, W4, mindisp, step, H4loc, q_row0, Hout, y_row0);
  else regress_fwd_kernel<16><<<grid, block, 0, st>>>(cost, disp, prob, H4glob, W4, mindisp, step, H4loc, q_row0, Hout, y_row0);
  return dpf::after_launch("dpf_regress_fwd");
}

extern "C" int dpf_regress_fwd(const float* cost, float* disp, float* prob, int B, int D, int H4, int W4, float mindisp,
                               float step, void* stream) {
  return dpf_regress_fwd_tile(cost, disp, prob, B, D, H4, W4, H4, 0, 4 * H4, 0, mindisp, step, stream);
}

extern "C" int dpf_regress_fwd_halfpixel(const float* cost, float* disp, float* prob, int B, int D, int H4, int W4, float mindisp,
                                         float step, void* stream) {
  DPF_REQUIRE(cost && disp, "dpf_regress_fwd_halfpixel: null pointer");
  DPF_REQUIRE(B > 0 && B <= 65535 && H4 > 1 && W4 > 1, "dpf_regress_fwd_halfpixel: bad shape B=%d H4=%d W4=%d", B, H4, W4);
  DPF_REQUIRE(D == 8, "dpf_regress_fwd_halfpixel: D=%d (only 8 levels)", D);
  dim3 grid((4 * W4 + kTX - 1) / kTX, (4 * H4 + kTY - 1) / kTY, B), block(kTX, kTY);
  regress_fwd_kernel<8, true><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(cost, disp, prob, H4, W4, mindisp, step, H4, 0,
                                                                                      4 * H4, 0);
  return dpf::after_launch("dpf_regress_fwd_halfpixel");
}

extern "C" int dpf_regress_bwd(const float* cost, const float* ddisp, float* dcost, int B, int D, int H4, int W4,
                               float mindisp, float step, void* stream) {
  DPF_REQUIRE(cost && ddisp && dcost, "dpf_regress_bwd: null pointer");
  DPF_REQUIRE(B > 0 && B <= 65535 && H4 > 1 && W4 > 1, "dpf_regress_bwd: bad shape B=%d H4=%d W4=%d", B, H4, W4);
  DPF_REQUIRE(D == 8 || D == 4 || D == 16, "dpf_regress_bwd: D=%d not in {4,8,16}", D);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(dcost, 0, static_cast<size_t>(B) * D * H4 * W4 * sizeof(float), st);
  if (e != cudaSuccess) return dpf::fail("dpf_regress_bwd: memset: %s", cudaGetErrorString(e));
  dim3 grid((4 * W4 + kTX - 1) / kTX, (4 * H4 + kTY - 1) / kTY, B), block(kTX, kTY);
  if (D == 8) regress_bwd_kernel<8><<<grid, block, 0, st>>>(cost, ddisp, dcost, H4, W4, mindisp, step);
  else if (D == 4) regress_bwd_kernel<4><<<grid, block, 0, st>>>(cost, ddisp, dcost, H4, W4, mindisp, step);
  else regress_bwd_kernel<16><<<grid, block, 0, st>>>(cost, ddisp, dcost, H4, W4, mindisp, step);
  return dpf::after_launch("dpf_regress_bwd");
}

extern "C" int dpf_softargmin_fwd(const float* cost, float* disp, float* prob, int B, int D, long long P, float mindisp, float step,
                                  void* stream) {
  DPF_REQUIRE(cost && disp, "dpf_softargmin_fwd: null pointer");
  DPF_REQUIRE(B > 0 && B <= 65535 && D >= 1 && D <= 64 && P > 0 && (P + 255) / 256 < (1LL << 31), "dpf_softargmin_fwd: bad shape");
  dim3 grid(static_cast<unsigned>((P + 255) / 256), B);
  softargmin_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(cost, disp, prob, D, P, mindisp, step);
  return dpf::after_launch("dpf_softargmin_fwd");
}
