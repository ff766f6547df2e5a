// 3-D deformable convolution (D3D) forward for sm_100a: trilinear gather fused into the A-operand producer of a
// tcgen05 GEMM -- no [27*Cin, B*D*H*W] column buffer (7-13 GB fp32 in the reference at 1120x1680).
//
// Replaces DCN.deform_conv_forward (src/module/dcn3d/src/deform_conv.h:10-29 -> src/cuda/deform_conv_cuda.cu:18-126)
// and its im2col kernel (src/cuda/deform_im2col_cuda.cuh:192-265, sampling rule :26-72, in-bounds test :248).
//   y[v, o] = relu?( scale[o] * sum_{tap,c} W[o,c,tap] * trilinear(x[:, c], p_v + tap - 1 + offset[v, 3*tap + (0,1,2)]) + shift[o] )
// x [B,D,H,W,x_cstride] bf16 (the first CINP channels are gathered; zero-padded beyond the real Cin), offset [B,D,H,W,81] fp32 ((d,h,w) per tap), y [B,D,H,W,64] bf16.
//
// Work unit: 256 consecutive voxels (2 GEMM blocks of 128 rows).  For every tap, 16 producer warps compute the
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
constexpr int kProdWarps = 16;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;   // 672
constexpr int kVoxPerPass = kProdWarps * 32 / 8;              // 8 threads (one 16-byte channel chunk each) per voxel
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
  static constexpr int A_CHUNK_BYTES = 128 * 16 + 16;                // chunk pitch (= LBO), +16 B: conflict-free st.shared
  static constexpr int A_BLOCK_BYTES = NCH * A_CHUNK_BYTES;          // [chunk][128 rows][16 B]
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
      mbar_init(&bar_full[i], kProdWarps + 1);       // gather warps + the weight-copy issuer (expect_tx)
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
    const int ptid = threadIdx.x - (kMmaWarp + 1) * 32;      // 0..511
    const int c8 = ptid & 7;
    const int vsub = ptid >> 3;                              // 0..63
    constexpr int PASSES = kBlocks * 128 / kVoxPerPass;
    const int HW = H * W;
    const int cs = p.x_cstride;
    const __nv_bfloat16* xc = p.x + (c8 < C::NCH ? c8 : 0) * 8;
    uint32_t g = 0;
    for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
      const long long v0 = static_cast<long long>(unit) * (kBlocks * 128);
      // voxel coordinates of this thread's PASSES rows: decomposed once per unit, reused by all 27 taps
      int vd[PASSES], vh[PASSES], vw[PASSES], vbase[PASSES];
      bool vlive[PASSES];
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const long long v = v0 + ps * kVoxPerPass + vsub;
        vlive[ps] = (v < p.nvox) && (c8 < C::NCH);
        int t = static_cast<int>(vlive[ps] ? v : 0);               // nvox < 2^31 (checked on the host)
        vw[ps] = t % W; t /= W;
        vh[ps] = t % H; t /= H;
        vd[ps] = t % D;
        vbase[ps] = (t / D) * D * HW;                               // voxel index of (b, 0, 0, 0)
      }
      for (int tap = 0; tap < kTaps; ++tap, ++g) {
        const int stage = g % kStages;
        const uint32_t ph = (g / kStages) & 1u;
        mbar_wait(&bar_empty[stage], ph ^ 1u);
        uint8_t* sa = s_stage + stage * C::STAGE_BYTES;
        if (ptid == 0) {
          mbar_arrive_expect_tx(&bar_full[stage], C::W_TAP_BYTES);
          bulk_g2s(smem_u32(sa + C::A_STAGE_BYTES), p.w + static_cast<size_t>(tap) * (C::W_TAP_BYTES / 2), C::W_TAP_BYTES, &bar_full[stage]);
        }
        const int ti = tap / 9 - 1, tj = (tap / 3) % 3 - 1, tk = tap % 3 - 1;
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
          const int r = ps * kVoxPerPass + vsub;             // row inside the work unit (0..255)
          const int vox = vbase[ps] + (vd[ps] * H + vh[ps]) * W + vw[ps];
          const float* op = p.offset + static_cast<size_t>(vox) * 81 + tap * 3;
          const float pd = static_cast<float>(vd[ps] + ti) + __ldg(op + 0);
          const float phh = static_cast<float>(vh[ps] + tj) + __ldg(op + 1);
          const float pw = static_cast<float>(vw[ps] + tk) + __ldg(op + 2);
          const bool inside = vlive[ps] && pd > -1.f && phh > -1.f && pw > -1.f && pd < static_cast<float>(D) &&
                              phh < static_cast<float>(H) && pw < static_cast<float>(W);
          const float fd = floorf(pd), fh = floorf(phh), fw = floorf(pw);
          const int d0 = static_cast<int>(fd), h0 = static_cast<int>(fh), w0 = static_cast<int>(fw);
          const float ld = pd - fd, lh = phh - fh, lw = pw - fw;
          // branch-free corners: clamp the index, zero the weight when the corner (or the whole sample) is outside,
          // so that all 8 loads are issued back to back
          const float wd0 = (inside && d0 >= 0) ? 1.f - ld : 0.f, wd1 = (inside && d0 + 1 <= D - 1) ? ld : 0.f;
          const float wh0 = (h0 >= 0) ? 1.f - lh : 0.f, wh1 = (h0 + 1 <= H - 1) ? lh : 0.f;
          const float ww0 = (w0 >= 0) ? 1.f - lw : 0.f, ww1 = (w0 + 1 <= W - 1) ? lw : 0.f;
          const int dc0 = min(max(d0, 0), D - 1), dc1 = min(max(d0 + 1, 0), D - 1);
          const int hc0 = min(max(h0, 0), H - 1), hc1 = min(max(h0 + 1, 0), H - 1);
          const int wc0 = min(max(w0, 0), W - 1), wc1 = min(max(w0 + 1, 0), W - 1);
          const int r00 = vbase[ps] + dc0 * HW + hc0 * W, r01 = vbase[ps] + dc0 * HW + hc1 * W;
          const int r10 = vbase[ps] + dc1 * HW + hc0 * W, r11 = vbase[ps] + dc1 * HW + hc1 * W;
          uint4 u[8];
          u[0] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<size_t>(r00 + wc0) * cs));
          u[1] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<size_t>(r00 + wc1) * cs));
          u[2] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<size_t>(r01 + wc0) * cs));
          u[3] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<size_t>(r01 + wc1) * cs));
          u[4] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<size_t>(r10 + wc0) * cs));
          u[5] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<size_t>(r10 + wc1) * cs));
          u[6] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<size_t>(r11 + wc0) * cs));
          u[7] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<size_t>(r11 + wc1) * cs));
          const float a00 = wd0 * wh0, a01 = wd0 * wh1, a10 = wd1 * wh0, a11 = wd1 * wh1;
          const float cw[8] = {a00 * ww0, a00 * ww1, a01 * ww0, a01 * ww1, a10 * ww0, a10 * ww1, a11 * ww0, a11 * ww1};
          // packed bf16 blend (HFMA2.BF16): the blended A tile is rounded to bf16 for the MMA anyway
          __nv_bfloat162 acc[4];
          {
            const __nv_bfloat162 w2 = __float2bfloat162_rn(cw[0]);
            acc[0] = __hmul2(w2, *reinterpret_cast<const __nv_bfloat162*>(&u[0].x));
            acc[1] = __hmul2(w2, *reinterpret_cast<const __nv_bfloat162*>(&u[0].y));
            acc[2] = __hmul2(w2, *reinterpret_cast<const __nv_bfloat162*>(&u[0].z));
            acc[3] = __hmul2(w2, *reinterpret_cast<const __nv_bfloat162*>(&u[0].w));
          }
#pragma unroll
          for (int c = 1; c < 8; ++c) {
            const __nv_bfloat162 w2 = __float2bfloat162_rn(cw[c]);
            acc[0] = __hfma2(w2, *reinterpret_cast<const __nv_bfloat162*>(&u[c].x), acc[0]);
            acc[1] = __hfma2(w2, *reinterpret_cast<const __nv_bfloat162*>(&u[c].y), acc[1]);
            acc[2] = __hfma2(w2, *reinterpret_cast<const __nv_bfloat162*>(&u[c].z), acc[2]);
            acc[3] = __hfma2(w2, *reinterpret_cast<const __nv_bfloat162*>(&u[c].w), acc[3]);
          }
          if (c8 < C::NCH) {
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&acc[0]); o.y = *reinterpret_cast<uint32_t*>(&acc[1]);
            o.z = *reinterpret_cast<uint32_t*>(&acc[2]); o.w = *reinterpret_cast<uint32_t*>(&acc[3]);
            const int blk = r >> 7, row = r & 127;
            *reinterpret_cast<uint4*>(sa + blk * C::A_BLOCK_BYTES + c8 * C::A_CHUNK_BYTES + row * 16) = o;
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
    const uint64_t adesc_hi = umma_desc_nosw(0, C::A_CHUNK_BYTES, 128);   // LBO = chunk pitch, SBO = 8 rows
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
              const uint64_t adesc = adesc_hi | static_cast<uint64_t>((a0 + blk * (C::A_BLOCK_BYTES >> 4) + ks * 2 * (C::A_CHUNK_BYTES >> 4)) & 0x3FFF);
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
  DPF_REQUIRE(static_cast<long long>(B) * D * H * W < (1LL << 31) / 128, "dpf_dcn3d_fwd: tensor too large for 32-bit voxel indexing");
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
