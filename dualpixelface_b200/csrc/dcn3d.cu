// 3-D deformable convolution (D3D) forward for sm_100a: trilinear gather fused into the A-operand producer of a
// tcgen05 GEMM -- no [27*Cin, B*D*H*W] column buffer (7-13 GB fp32 in the reference at 1120x1680).
//
// Replaces DCN.deform_conv_forward (src/module/dcn3d/src/deform_conv.h:10-29 -> src/cuda/deform_conv_cuda.cu:18-126)
// and its im2col kernel (src/cuda/deform_im2col_cuda.cuh:192-265, sampling rule :26-72, in-bounds test :248).
//   y[v, o] = relu?( scale[o] * sum_{tap,c} W[o,c,tap] * trilinear(x[:, c], p_v + tap - 1 + offset[v, 3*tap + (0,1,2)]) + shift[o] )
// x [B,D,H,W,x_cstride] bf16 (the first CINP channels are gathered; zero-padded beyond the real Cin), offset [B,D,H,W,81] fp32 ((d,h,w) per tap), y [B,D,H,W,64] bf16.
//
// Work unit: 256 consecutive voxels (2 GEMM blocks of 128 rows).  For every tap, 8 producer warps compute the
// trilinear sample of all CINP channels (8 threads per voxel, one 16-byte channel chunk each; fp32 blend) and write it
// as the bf16 A tile in the UMMA no-swizzle K-major layout; the tap's weight tile [CINP x 64] is streamed next to it by a
// 1-D TMA bulk copy.  One elected lane issues the tcgen05.mma; accumulators are double-buffered in TMEM so the
// epilogue of a unit overlaps the gather of the next.
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"

namespace {

using namespace dpf;

constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kProdWarps = 8;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;   // 416
constexpr int kNOut = 64;
constexpr int kBlocks = 2;                                    // GEMM blocks (128 voxels each) per work unit
constexpr int kStages = 3;
constexpr int kTaps = 27;

struct DcnParams {
  const __nv_bfloat16* x;
  const float* offset;
  const __nv_bfloat16* w;
  const float* scale;
  const float* shift;
  __nv_bfloat16* y;
  int B, D, H, W, relu, x_cstride;
  long long nvox;
  int nunits;
};

template <int CINP>
struct DCfg {
  static constexpr int NCH = CINP / 8;
  static constexpr int KSTEPS = CINP / 16;
  static constexpr int A_BLOCK_BYTES = NCH * 128 * 16;               // [chunk][128 rows][16 B]
  static constexpr int A_STAGE_BYTES = kBlocks * A_BLOCK_BYTES;
  static constexpr int W_TAP_BYTES = NCH * kNOut * 16;               // [chunk][64 rows][16 B]
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + W_TAP_BYTES;
  static constexpr int TMEM_COLS = 256;                              // 2 stages x 2 blocks x 64 columns
  static constexpr int SMEM_BYTES = kStages * STAGE_BYTES + 2 * kNOut * 4 + (2 * kStages + 4) * 8 + 16 + 128;
  static_assert(CINP % 16 == 0, "CINP must be a multiple of 16");
};

template <int CINP>
__global__ void __launch_bounds__(kThreads, 1) dcn3d_kernel(const __grid_constant__ DcnParams p) {
  using C = DCfg<CINP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  uint8_t* s_stage = smem;
  float* s_scale = reinterpret_cast<float*>(smem + kStages * C::STAGE_BYTES);
  float* s_shift = s_scale + kNOut;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_shift + kNOut);
  uint64_t* bar_empty = bar_full + kStages;
  uint64_t* bar_tfull = bar_empty + kStages;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < kNOut; i += kThreads) {
    s_scale[i] = p.scale ? p.scale[i] : 1.0f;
    s_shift[i] = p.shift ? p.shift[i] : 0.0f;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bar_full[i], kProdWarps + 1);       // 8 gather warps + the weight-copy issuer (expect_tx)
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(s_tmem, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const int D = p.D, H = p.H, W = p.W;

  if (warp > kMmaWarp) {
    // ======================= producers: trilinear gather -> bf16 A tile; weight tile by TMA ==================
    const int ptid = threadIdx.x - (kMmaWarp + 1) * 32;      // 0..255
    const int c8 = ptid & 7;
    const int vsub = ptid >> 3;                              // 0..31
    uint32_t g = 0;
    for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
      const long long v0 = static_cast<long long>(unit) * (kBlocks * 128);
      for (int tap = 0; tap < kTaps; ++tap, ++g) {
        const int stage = g % kStages;
        const uint32_t ph = (g / kStages) & 1u;
        mbar_wait(&bar_empty[stage], ph ^ 1u);
        uint8_t* sa = s_stage + stage * C::STAGE_BYTES;
        if (ptid == 0) {
          mbar_arrive_expect_tx(&bar_full[stage], C::W_TAP_BYTES);
          bulk_g2s(smem_u32(sa + C::A_STAGE_BYTES), p.w + static_cast<size_t>(tap) * (C::W_TAP_BYTES / 2), C::W_TAP_BYTES, &bar_full[stage]);
        }
        const int ti = tap / 9, tj = (tap / 3) % 3, tk = tap % 3;
#pragma unroll 2
        for (int pass = 0; pass < kBlocks * 128 / 32; ++pass) {
          const int r = pass * 32 + vsub;                    // row inside the work unit (0..255)
          const long long v = v0 + r;
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (v < p.nvox && c8 < C::NCH) {
            long long t = v;
            const int w = static_cast<int>(t % W); t /= W;
            const int h = static_cast<int>(t % H); t /= H;
            const int d = static_cast<int>(t % D);
            const int b = static_cast<int>(t / D);
            const float* op = p.offset + v * 81 + tap * 3;
            const float pd = static_cast<float>(d - 1 + ti) + __ldg(op + 0);
            const float phh = static_cast<float>(h - 1 + tj) + __ldg(op + 1);
            const float pw = static_cast<float>(w - 1 + tk) + __ldg(op + 2);
            if (pd > -1.f && phh > -1.f && pw > -1.f && pd < static_cast<float>(D) && phh < static_cast<float>(H) && pw < static_cast<float>(W)) {
              const int d0 = static_cast<int>(floorf(pd)), h0 = static_cast<int>(floorf(phh)), w0 = static_cast<int>(floorf(pw));
              const float ld = pd - static_cast<float>(d0), lh = phh - static_cast<float>(h0), lw = pw - static_cast<float>(w0);
              const __nv_bfloat16* xb = p.x + static_cast<size_t>(b) * D * H * W * p.x_cstride + c8 * 8;
#pragma unroll
              for (int cd = 0; cd < 2; ++cd) {
                const int di = d0 + cd;
                if (di < 0 || di > D - 1) continue;
                const float wd = cd ? ld : 1.f - ld;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                  const int hi = h0 + ch;
                  if (hi < 0 || hi > H - 1) continue;
                  const float wdh = wd * (ch ? lh : 1.f - lh);
#pragma unroll
                  for (int cw = 0; cw < 2; ++cw) {
                    const int wi = w0 + cw;
                    if (wi < 0 || wi > W - 1) continue;
                    const float ww = wdh * (cw ? lw : 1.f - lw);
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(xb + ((static_cast<size_t>(di) * H + hi) * W + wi) * p.x_cstride));
                    acc[0] = fmaf(ww, bf16_lo(u.x), acc[0]); acc[1] = fmaf(ww, bf16_hi(u.x), acc[1]);
                    acc[2] = fmaf(ww, bf16_lo(u.y), acc[2]); acc[3] = fmaf(ww, bf16_hi(u.y), acc[3]);
                    acc[4] = fmaf(ww, bf16_lo(u.z), acc[4]); acc[5] = fmaf(ww, bf16_hi(u.z), acc[5]);
                    acc[6] = fmaf(ww, bf16_lo(u.w), acc[6]); acc[7] = fmaf(ww, bf16_hi(u.w), acc[7]);
                  }
                }
              }
            }
          }
          if (c8 < C::NCH) {
            uint4 o;
            o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
            o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
            const int blk = r >> 7, row = r & 127;
            *reinterpret_cast<uint4*>(sa + blk * C::A_BLOCK_BYTES + (c8 * 128 + row) * 16) = o;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[stage]);
      }
    }
  } else if (warp == kMmaWarp) {
    // ======================================= MMA issuer =====================================================
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, kNOut);
    const uint64_t adesc_hi = umma_desc_nosw(0, 128 * 16, 128);        // LBO = chunk pitch (2 KB), SBO = 8 rows
    const uint64_t bdesc_hi = umma_desc_nosw(0, kNOut * 16, 128);
    const uint32_t sbase = smem_u32(s_stage);
    const bool leader = elect_one();
    uint32_t g = 0, it = 0;
    for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x, ++it) {
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      mbar_wait(&bar_tempty[as], aph ^ 1u);
      tc_fence_after_sync();
      for (int tap = 0; tap < kTaps; ++tap, ++g) {
        const int stage = g % kStages;
        mbar_wait(&bar_full[stage], (g / kStages) & 1u);
        tc_fence_after_sync();
        if (leader) {
          const uint32_t a0 = (sbase + stage * C::STAGE_BYTES) >> 4;
          const uint32_t b0 = (sbase + stage * C::STAGE_BYTES + C::A_STAGE_BYTES) >> 4;
#pragma unroll
          for (int ks = 0; ks < C::KSTEPS; ++ks) {
            const uint64_t bdesc = bdesc_hi | static_cast<uint64_t>((b0 + ks * 2 * kNOut) & 0x3FFF);
#pragma unroll
            for (int blk = 0; blk < kBlocks; ++blk) {
              const uint64_t adesc = adesc_hi | static_cast<uint64_t>((a0 + blk * (C::A_BLOCK_BYTES >> 4) + ks * 2 * 128) & 0x3FFF);
              umma_bf16(tmem_base + (as * kBlocks + blk) * kNOut, adesc, bdesc, idesc, !(tap == 0 && ks == 0));
            }
          }
          umma_commit(&bar_empty[stage]);
          if (tap == kTaps - 1) umma_commit(&bar_tfull[as]);
        }
        __syncwarp();
      }
    }
  } else {
    // ======================================= epilogue =======================================================
    uint32_t it = 0;
    for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x, ++it) {
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      mbar_wait(&bar_tfull[as], aph);
      tc_fence_after_sync();
#pragma unroll 1
      for (int blk = 0; blk < kBlocks; ++blk) {
        const long long v = static_cast<long long>(unit) * (kBlocks * 128) + blk * 128 + warp * 32 + lane;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + (as * kBlocks + blk) * kNOut;
#pragma unroll
        for (int c0 = 0; c0 < kNOut; c0 += 16) {
          uint32_t r[16];
          __syncwarp();
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
          if (v >= p.nvox) continue;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            f[j] = __uint_as_float(r[j]) * s_scale[c0 + j] + s_shift[c0 + j];
            if (p.relu) f[j] = fmaxf(f[j], 0.f);
          }
          uint4 o0, o1;
          o0.x = pack_bf16x2(f[0], f[1]); o0.y = pack_bf16x2(f[2], f[3]); o0.z = pack_bf16x2(f[4], f[5]); o0.w = pack_bf16x2(f[6], f[7]);
          o1.x = pack_bf16x2(f[8], f[9]); o1.y = pack_bf16x2(f[10], f[11]); o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
          uint4* dst = reinterpret_cast<uint4*>(p.y + v * kNOut + c0);
          dst[0] = o0;
          dst[1] = o1;
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[as]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int CINP>
int launch_dcn(const DcnParams& p, cudaStream_t st) {
  using C = DCfg<CINP>;
  auto kern = dcn3d_kernel<CINP>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return dpf::fail("dpf_dcn3d_fwd: cannot opt in to %d B shared memory: %s", C::SMEM_BYTES, cudaGetErrorString(e));
    attr_done = true;
  }
  const int grid = std::min(p.nunits, dpf::sm_count());
  kern<<<grid, kThreads, C::SMEM_BYTES, st>>>(p);
  return dpf::after_launch("dpf_dcn3d_fwd");
}

}  // namespace

extern "C" int dpf_dcn3d_fwd(const void* x, const float* offset, const void* w, const float* scale, const float* shift,
                             void* y, int B, int D, int H, int W, int Cin_pad, int x_cstride, int Cout, int relu, void* stream) {
  DPF_REQUIRE(x && offset && w && y, "dpf_dcn3d_fwd: null pointer");
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(w) && DPF_ALIGNED16(y), "dpf_dcn3d_fwd: pointers must be 16-byte aligned");
  DPF_REQUIRE(Cout == kNOut, "dpf_dcn3d_fwd: Cout=%d, only 64 is built", Cout);
  DPF_REQUIRE(Cin_pad == 48 || Cin_pad == 64, "dpf_dcn3d_fwd: Cin_pad=%d must be 48 or 64", Cin_pad);
  DPF_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "dpf_dcn3d_fwd: bad shape");
  DPF_REQUIRE(x_cstride >= Cin_pad && x_cstride % 8 == 0, "dpf_dcn3d_fwd: x_cstride=%d must be a multiple of 8 >= Cin_pad", x_cstride);
  DcnParams p{};
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.offset = offset;
  p.w = reinterpret_cast<const __nv_bfloat16*>(w);
  p.scale = scale;
  p.shift = shift;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.B = B; p.D = D; p.H = H; p.W = W; p.relu = relu; p.x_cstride = x_cstride;
  p.nvox = static_cast<long long>(B) * D * H * W;
  p.nunits = static_cast<int>((p.nvox + kBlocks * 128 - 1) / (kBlocks * 128));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (Cin_pad == 48) return launch_dcn<48>(p, st);
  return launch_dcn<64>(p, st);
}
