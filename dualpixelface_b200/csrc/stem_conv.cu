// Encoder stem: 3x3 / stride 2 / pad 1 convolution from an 8-channel (3 real + 5 zero) channels-last bf16 image to 32 channels,
// folded BatchNorm bias + ReLU in the epilogue (sm_100a; adjacent to the hot path).
//
// Replaces `convbn(input_channel, 32, 3, 2, 1, 1)` + ReLU, the first layer of both encoders
// (src/model/stereodpnet/modules.py:66-68, src/model/psmnet/modules.py:72-74; convbn = src/module/asm/basics.py:17-22).
// cuDNN runs this layer (K = 27) on an sm80-class kernel at 0.40 ms for 0.07 ms of HBM traffic (8 x 1120 x 1680 -> 560 x 840).
//
// The layer is pure bandwidth (482 MB moved, 6.5 GFLOP useful), so it runs on warp-level mma.sync (m16n8k16, bf16 -> fp32) rather
// than on the tcgen05 pipeline: a CTA stages the (2*8+1) x (2*64+1) input window of an 8 x 64 output tile once (16-byte cp.async,
// zero fill = the padding); GEMM K = tap * 8 + ci (9 taps x 8 channels = 72, padded to 80 = 5 k-steps), so an A-fragment register
// is ONE 32-bit shared-memory word (two channels of one tap of one pixel) -- no im2col buffer; the 80 x 32 weight matrix lives in
// registers as B fragments; results are staged in shared memory (80-byte pixel pitch, conflict-free) and leave as full 64-byte
// pixel rows with 16-byte coalesced stores.
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"
#include <algorithm>

namespace {

using namespace dpf;

constexpr int OTH = 8, OTW = 64;                 // output tile (one warp per output row)
constexpr int ITH = 2 * OTH + 1, ITW = 2 * OTW + 1;
constexpr int kThreads = OTH * 32;
constexpr int KSTEPS = 5;                        // K = 80 = 10 tap slots x 8 channels (tap 9 and channels 3..7 carry zero weights)
constexpr int IN_BYTES = ITH * ITW * 16;
constexpr int OUT_PITCH = 80;                    // bytes per staged output pixel (64 + 16: 8 pixels x 4 lanes hit 32 distinct banks)
constexpr int OUT_BYTES = OTH * OTW * OUT_PITCH;
constexpr int SMEM_BYTES = IN_BYTES + OUT_BYTES;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// x [N,H,W,8] bf16, w [80][32] bf16 (k = tap*8 + ci, row-major), bias fp32 [32], y [N,Ho,Wo,32] bf16
__global__ void __launch_bounds__(kThreads) stem_conv_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                             const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int H, int W,
                                                             int Ho, int Wo, int relu) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* s_in = smem;
  uint8_t* s_out = smem + IN_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n = blockIdx.z;
  const int oy0 = blockIdx.y * OTH, ox0 = blockIdx.x * OTW;
  const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;

  // ---- stage the input window (zero fill outside the image = the convolution padding)
  const __nv_bfloat16* xn = x + static_cast<size_t>(n) * H * W * 8;
  for (int i = threadIdx.x; i < ITH * ITW; i += kThreads) {
    const int r = i / ITW, c = i - r * ITW;
    const int iy = iy0 + r, ix = ix0 + c;
    const bool ok = (iy >= 0) && (iy < H) && (ix >= 0) && (ix < W);
    cp_async16_zfill(smem_u32(s_in + i * 16), ok ? xn + (static_cast<size_t>(iy) * W + ix) * 8 : x, ok);
  }
  cp_async_commit();

  // ---- B fragments of the whole 80 x 32 weight matrix: b0/b1 = W[16 s + 2 t (+1)][8 j + g], b2/b3 = rows + 8
  uint32_t bw[KSTEPS][4][2];
#pragma unroll
  for (int s = 0; s < KSTEPS; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = 8 * j + g;
      const int k0 = 16 * s + 2 * t;
      const uint32_t lo0 = __bfloat16_as_ushort(w[(k0)*32 + col]), hi0 = __bfloat16_as_ushort(w[(k0 + 1) * 32 + col]);
      const uint32_t lo1 = __bfloat16_as_ushort(w[(k0 + 8) * 32 + col]), hi1 = __bfloat16_as_ushort(w[(k0 + 9) * 32 + col]);
      bw[s][j][0] = lo0 | (hi0 << 16);
      bw[s][j][1] = lo1 | (hi1 << 16);
    }
  float bia[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bia[j][0] = bias ? __ldg(bias + 8 * j + 2 * t) : 0.f;
    bia[j][1] = bias ? __ldg(bias + 8 * j + 2 * t + 1) : 0.f;
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- one warp = one output row of the tile: 4 m16 tiles of 16 pixels
  const int oyl = warp;
#pragma unroll 1
  for (int mt = 0; mt < OTW / 16; ++mt) {
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const int px0 = mt * 16 + g, px1 = px0 + 8;               // output columns (tile-local) of this thread's two fragment rows
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s) {
      uint32_t a[4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int tap = min(2 * s + half, 8);                  // tap slot 9 has zero weights: read tap 8 again (any valid word)
        const int kh = tap / 3, kw = tap - 3 * kh;
        const uint8_t* row = s_in + ((2 * oyl + kh) * ITW + kw) * 16 + 4 * t;
        a[2 * half + 0] = *reinterpret_cast<const uint32_t*>(row + (2 * px0) * 16);
        a[2 * half + 1] = *reinterpret_cast<const uint32_t*>(row + (2 * px1) * 16);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_bf16_16816(acc[j], a, bw[s][j][0], bw[s][j][1]);
    }
    // ---- bias + ReLU -> staged bf16 (thread: rows px0 / px1, channels 8 j + 2 t, +1)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v0 = acc[j][0] + bia[j][0], v1 = acc[j][1] + bia[j][1], v2 = acc[j][2] + bia[j][0], v3 = acc[j][3] + bia[j][1];
      if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
      uint8_t* o = s_out + (oyl * OTW) * OUT_PITCH + (8 * j + 2 * t) * 2;
      *reinterpret_cast<uint32_t*>(o + px0 * OUT_PITCH) = pack_bf16x2(v0, v1);
      *reinterpret_cast<uint32_t*>(o + px1 * OUT_PITCH) = pack_bf16x2(v2, v3);
    }
  }
  __syncthreads();

  // ---- coalesced write-out: 4 x 16-byte pieces per pixel, consecutive threads -> consecutive pieces
  __nv_bfloat16* yn = y + static_cast<size_t>(n) * Ho * Wo * 32;
  for (int i = threadIdx.x; i < OTH * OTW * 4; i += kThreads) {
    const int px = i >> 2, piece = i & 3;
    const int r = px / OTW, c = px - r * OTW;
    const int oy = oy0 + r, ox = ox0 + c;
    if (oy < Ho && ox < Wo) {
      const uint4 v = *reinterpret_cast<const uint4*>(s_out + px * OUT_PITCH + piece * 16);
      *reinterpret_cast<uint4*>(yn + (static_cast<size_t>(oy) * Wo + ox) * 32 + piece * 8) = v;
    }
  }
}

}  // namespace

extern "C" int dpf_stem_conv_fwd(const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int relu, void* stream) {
  DPF_REQUIRE(x && w && y, "dpf_stem_conv_fwd: null pointer");
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(y), "dpf_stem_conv_fwd: x and y must be 16-byte aligned");
  DPF_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0, "dpf_stem_conv_fwd: bad shape");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return dpf::fail("dpf_stem_conv_fwd: cannot opt in to %d B shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
    attr_done = true;
  }
  dim3 grid((Wo + OTW - 1) / OTW, (Ho + OTH - 1) / OTH, N);
  DPF_REQUIRE(grid.y <= 65535, "dpf_stem_conv_fwd: image too tall");
  stem_conv_kernel<<<grid, kThreads, SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(w), bias, reinterpret_cast<__nv_bfloat16*>(y), H, W,
      Ho, Wo, relu);
  return dpf::after_launch("dpf_stem_conv_fwd");
}
