// Backward of the 3-D deformable convolution (D3D) for sm_100a.
//
// Replaces DCN.deform_conv_backward of the reference (src/module/dcn3d/src/cuda/deform_conv_cuda.cu:128-285 and the
// col2im / col2im_coord kernels of src/cuda/deform_im2col_cuda.cuh:267-405), again without any column buffer:
//
//   dpf_dcn3d_bwd_data   dcol[v, c] = sum_o dy[v, o] * W[o, c, tap] per tap on tcgen05 (A = the dy tile, B = W_tap^T streamed by
//                        TMA bulk copies, accumulators in TMEM).  The epilogue threads (one per voxel) then
//                          - scatter  dx[corner, c] += w_corner * dcol[c]           (vector fp32 atomics, like the reference)
//                          - compute  doffset[v, 3*tap + axis] = sum_c dcol[c] * d(sample_c)/d(axis)
//                            = sum over the 8 corners of (+-1)*(the two other axes' weights) * <dcol, x[corner]>
//                        with the reference's validity rules (deform_im2col_cuda.cuh:111-190,248,313-331).
//   dpf_dcn3d_bwd_weight dW[tap][c][o] = sum_v col[v, tap, c] * dy[v, o]: the forward's gather producer rebuilds the sampled
//                        A tile, which is then used as an MN-major operand with the voxels as the K dimension (M = 64
//                        input channels, N = 64 output channels, two taps interleaved per TMEM column block); a launch
//                        covers a range of <= 14 taps so that the accumulators fit the 512 TMEM columns.
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"

namespace {

using namespace dpf;

constexpr int kC = 64;                      // channels gathered (zero padded) and output channels
constexpr int kNCH = kC / 8;
constexpr int kChunk = 128 * 16 + 16;       // chunk-plane pitch of a [c8][128 rows][16 B] block (conflict-free stores)
constexpr int kBlock = kNCH * kChunk;       // one 128-row block
constexpr int kTile = 2 * kBlock;           // 256-voxel unit
constexpr int kWTap = kNCH * kC * 16;       // one packed weight tap [c8][64][16 B] = 8 KB
constexpr int kTaps = 27;

struct BwdParams {
  const __nv_bfloat16* x;
  const float* offset;
  const __nv_bfloat16* dy;
  const __nv_bfloat16* w;       // data kernel: W^T packed [tap][o/8][c][8]
  float* dx;                    // [nvox][x_cstride] fp32 (zeroed by the caller)
  float* doff;                  // [nvox][off_cstride] fp32 (channels >= 81 are not written)
  float* dw;                    // weight kernel: [27][c][o] fp32 (zeroed by the caller)
  int B, D, H, W, x_cstride, off_cstride;
  int nunits, tiles_h, tiles_w, tap0, tap1;
  int dx_channels;              // data kernel: 32 -> only channels [0,32) of dx are accumulated, else all 64
};

__device__ __forceinline__ void unit_coords(int unit, const BwdParams& p, int& d, int& h0, int& w0, int& b) {
  d = unit % p.D;
  int t = unit / p.D;
  w0 = (t % p.tiles_w) * 16;
  t /= p.tiles_w;
  h0 = (t % p.tiles_h) * 16;
  b = t / p.tiles_h;
}

// cp.async a [256 voxel x 64 channel] bf16 tile of a [nvox, 64] tensor into the chunk-planar layout (zero fill outside)
__device__ __forceinline__ void load_tile_256x64(uint8_t* dst, const __nv_bfloat16* src, const BwdParams& p, int ud, int uh0,
                                                 int uw0, int ub, int tid, int nthreads) {
  for (int q = tid; q < 256 * kNCH; q += nthreads) {
    const int r = q >> 3, c8 = q & 7;
    const int hh = uh0 + (r >> 4), ww = uw0 + (r & 15);
    const bool ok = hh < p.H && ww < p.W;
    const size_t v = ((static_cast<size_t>(ub) * p.D + ud) * p.H + (ok ? hh : 0)) * p.W + (ok ? ww : 0);
    cp_async16_zfill(smem_u32(dst + (r >> 7) * kBlock + c8 * kChunk + (r & 127) * 16), src + v * kC + c8 * 8, ok);
  }
}

// ============================================================================================================
// data + offset gradient
// ============================================================================================================
constexpr int kDEpi = 16;                         // epilogue warps: TMEM lane quarter = warp & 3, 128-row block = (warp >> 2) & 1,
                                                  // 16-row half of the quarter = warp >> 3
constexpr int kDMma = kDEpi;                      // MMA warp index
constexpr int kDThreads = (kDEpi + 1 + 2) * 32;   // + MMA warp + 2 loader warps
constexpr int kDWStages = 3;
constexpr int kGPitch = kC + 4;                   // floats per staged dcol row (conflict-free 128-bit row writes)
constexpr int kGTile = 16 * kGPitch * 4;          // one warp's [16 voxels][64 channels] fp32 staging tile
constexpr int kDSmem = 2 * kTile + kDWStages * kWTap + kDEpi * kGTile + (2 * 2 + 2 * kDWStages + 4) * 8 + 16 + 128;

__global__ void __launch_bounds__(kDThreads, 1) dcn3d_bwd_data_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  uint8_t* s_dy = smem;                                   // [2][kTile]
  uint8_t* s_w = smem + 2 * kTile;                        // [kDWStages][kWTap]
  float* s_g = reinterpret_cast<float*>(s_w + kDWStages * kWTap);   // [kDEpi][16][kGPitch]
  uint64_t* bar_zfull = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_g) + kDEpi * kGTile);
  uint64_t* bar_zempty = bar_zfull + 2;
  uint64_t* bar_wfull = bar_zempty + 2;
  uint64_t* bar_wempty = bar_wfull + kDWStages;
  uint64_t* bar_tfull = bar_wempty + kDWStages;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_zfull[i], 2); mbar_init(&bar_zempty[i], 1); mbar_init(&bar_tfull[i], 1); mbar_init(&bar_tempty[i], kDEpi); }
    for (int i = 0; i < kDWStages; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], 1); }
    mbar_fence_init();
  }
  if (warp == kDMma) { tmem_alloc(s_tmem, 256); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const int D = p.D, H = p.H, W = p.W, HW = p.H * p.W;
  const int per_cta = (p.nunits + gridDim.x - 1) / gridDim.x;
  const int unit_lo = min(static_cast<int>(blockIdx.x) * per_cta, p.nunits), unit_hi = min(unit_lo + per_cta, p.nunits);

  if (warp > kDMma) {
    // ---------------- loaders: dy tile per unit (cp.async), W^T tap tiles (TMA bulk) ----------------------
    const int ltid = threadIdx.x - (kDMma + 1) * 32;
    uint32_t gw = 0, it = 0;
    for (int unit = unit_lo; unit < unit_hi; ++unit, ++it) {
      int ud, uh0, uw0, ub;
      unit_coords(unit, p, ud, uh0, uw0, ub);
      mbar_wait(&bar_zempty[it & 1], ((it >> 1) & 1u) ^ 1u);
      load_tile_256x64(s_dy + (it & 1) * kTile, p.dy, p, ud, uh0, uw0, ub, ltid, 64);
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_zfull[it & 1]);
      if (ltid == 0) {
        for (int tap = 0; tap < kTaps; ++tap, ++gw) {
          const int ws = gw % kDWStages;
          mbar_wait(&bar_wempty[ws], ((gw / kDWStages) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&bar_wfull[ws], kWTap);
          bulk_g2s(smem_u32(s_w + ws * kWTap), p.w + static_cast<size_t>(tap) * (kWTap / 2), kWTap, &bar_wfull[ws]);
        }
      }
    }
  } else if (warp == kDMma) {
    // ---------------- MMA: dcol[128 x 64] = dy[128 x 64(o)] * W_tap^T[64(o) x 64(c)] ------------------------
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, kC);
    const uint64_t adesc_hi = umma_desc_nosw(0, kChunk, 128);
    const uint64_t bdesc_hi = umma_desc_nosw(0, kC * 16, 128);
    const bool leader = elect_one();
    uint32_t gw = 0, it = 0, ga = 0;
    for (int unit = unit_lo; unit < unit_hi; ++unit, ++it) {
      mbar_wait(&bar_zfull[it & 1], (it >> 1) & 1u);
      tc_fence_after_sync();
      const uint32_t a_tile = smem_u32(s_dy + (it & 1) * kTile) >> 4;
      for (int tap = 0; tap < kTaps; ++tap, ++gw, ++ga) {
        const int ws = gw % kDWStages;
        const uint32_t as = ga & 1u;
        mbar_wait(&bar_wfull[ws], (gw / kDWStages) & 1u);
        mbar_wait(&bar_tempty[as], ((ga >> 1) & 1u) ^ 1u);
        tc_fence_after_sync();
        if (leader) {
          const uint32_t b0 = smem_u32(s_w + ws * kWTap) >> 4;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bdesc = bdesc_hi | static_cast<uint64_t>((b0 + ks * 2 * kC) & 0x3FFF);
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
              const uint64_t adesc = adesc_hi | static_cast<uint64_t>((a_tile + blk * (kBlock >> 4) + ks * 2 * (kChunk >> 4)) & 0x3FFF);
              umma_bf16(tmem_base + (as * 2 + blk) * kC, adesc, bdesc, idesc, ks > 0);
            }
          }
          umma_commit(&bar_wempty[ws]);
          umma_commit(&bar_tfull[as]);
          if (tap == kTaps - 1) umma_commit(&bar_zempty[it & 1]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue: TMEM -> per-warp smem tile -> 8 lanes per voxel scatter dx / reduce doffset --
    // Lane j of a voxel's 8-lane group owns channels [4j, 4j+4) and [32+4j, 32+4j+4): each red.global.add.v4.f32 of a group
    // covers one full 128-byte line of dx, each 64-bit x load one 64-byte half line.
    const int quarter = warp & 3, blk = (warp >> 2) & 1, half = warp >> 3;
    float* g_tile = s_g + warp * (16 * kGPitch);
    const int grp = lane >> 3, j = lane & 7;
    const uint32_t cs = static_cast<uint32_t>(p.x_cstride);
    const bool hi_live = p.dx_channels > 32;
    uint32_t ga = 0;
    for (int unit = unit_lo; unit < unit_hi; ++unit) {
      int ud, uh0, uw0, ub;
      unit_coords(unit, p, ud, uh0, uw0, ub);
      const int vbase = ub * D * HW;
      for (int tap = 0; tap < kTaps; ++tap, ++ga) {
        const uint32_t as = ga & 1u;
        mbar_wait(&bar_tfull[as], (ga >> 1) & 1u);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (as * 2 + blk) * kC;
#pragma unroll
        for (int c0 = 0; c0 < kC; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if ((lane >> 4) == half) {                           // this warp scatters rows [16*half, 16*half+16) of the quarter
            float4* dst = reinterpret_cast<float4*>(g_tile + (lane & 15) * kGPitch + c0);
            dst[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
            dst[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
            dst[2] = make_float4(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
            dst[3] = make_float4(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]), __uint_as_float(v[15]));
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[as]);          // the accumulator is free again while this warp scatters
        const int ti = tap / 9 - 1, tj = (tap / 3) % 3 - 1, tk = tap % 3 - 1;
#pragma unroll 1
        for (int rnd = 0; rnd < 4; ++rnd) {
          const int row = rnd * 4 + grp;                       // row of this warp's 16-voxel slice
          const int r = blk * 128 + quarter * 32 + half * 16 + row;
          const int hh = uh0 + (r >> 4), ww = uw0 + (r & 15);
          const bool live = hh < H && ww < W;
          const int vox = vbase + (ud * H + (live ? hh : 0)) * W + (live ? ww : 0);
          const float* op = p.offset + static_cast<size_t>(vox) * p.off_cstride + tap * 3;
          const float pd = static_cast<float>(ud + ti) + __ldg(op + 0);
          const float ph = static_cast<float>(hh + tj) + __ldg(op + 1);
          const float pw = static_cast<float>(ww + tk) + __ldg(op + 2);
          const bool inside = live && pd > -1.f && ph > -1.f && pw > -1.f && pd < static_cast<float>(D) &&
                              ph < static_cast<float>(H) && pw < static_cast<float>(W);
          const float4 ga4 = *reinterpret_cast<const float4*>(g_tile + row * kGPitch + 4 * j);
          const float4 gb4 = *reinterpret_cast<const float4*>(g_tile + row * kGPitch + 32 + 4 * j);
          const float fd = floorf(pd), fh = floorf(ph), fw = floorf(pw);
          const int d0 = static_cast<int>(fd), h0 = static_cast<int>(fh), w0 = static_cast<int>(fw);
          const float ld = pd - fd, lh = ph - fh, lw = pw - fw;
          float gd = 0.f, gh = 0.f, gw2 = 0.f;                 // this lane's partial sums (its 8 channels); reduced once below
#pragma unroll
          for (int cd = 0; cd < 2; ++cd) {
            // the 8 loads of one depth pair of corners are issued before any use (branch-free clamped addresses)
            uint2 xa[4], xb[4];
            uint32_t cidx[4];
            const int dc = min(max(d0 + cd, 0), D - 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int hi = min(max(h0 + (q >> 1), 0), H - 1), wi = min(max(w0 + (q & 1), 0), W - 1);
              cidx[q] = static_cast<uint32_t>(vbase + dc * HW + hi * W + wi) * cs;
              xa[q] = __ldg(reinterpret_cast<const uint2*>(p.x + cidx[q] + 4 * j));
              xb[q] = __ldg(reinterpret_cast<const uint2*>(p.x + cidx[q] + 32 + 4 * j));
            }
            const float wd = cd ? ld : 1.f - ld;
            const bool dok = inside && d0 + cd >= 0 && d0 + cd <= D - 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int ch = q >> 1, cw = q & 1;
              const int hi = h0 + ch, wi = w0 + cw;
              const bool ok = dok && hi >= 0 && hi <= H - 1 && wi >= 0 && wi <= W - 1;
              const float wh = ch ? lh : 1.f - lh, wv = cw ? lw : 1.f - lw;
              float s = ga4.x * bf16_lo(xa[q].x) + ga4.y * bf16_hi(xa[q].x) + ga4.z * bf16_lo(xa[q].y) + ga4.w * bf16_hi(xa[q].y) +
                        gb4.x * bf16_lo(xb[q].x) + gb4.y * bf16_hi(xb[q].x) + gb4.z * bf16_lo(xb[q].y) + gb4.w * bf16_hi(xb[q].y);
              s = ok ? s : 0.f;
              if (ok) {
                const float wc = wd * wh * wv;
                float* dxp = p.dx + cidx[q];
                atomicAdd(reinterpret_cast<float4*>(dxp + 4 * j), make_float4(wc * ga4.x, wc * ga4.y, wc * ga4.z, wc * ga4.w));
                if (hi_live) atomicAdd(reinterpret_cast<float4*>(dxp + 32 + 4 * j), make_float4(wc * gb4.x, wc * gb4.y, wc * gb4.z, wc * gb4.w));
              }
              gd += (cd ? 1.f : -1.f) * wh * wv * s;
              gh += (ch ? 1.f : -1.f) * wd * wv * s;
              gw2 += (cw ? 1.f : -1.f) * wd * wh * s;
            }
          }
#pragma unroll
          for (int m = 1; m < 8; m <<= 1) {
            gd += __shfl_xor_sync(0xffffffffu, gd, m);
            gh += __shfl_xor_sync(0xffffffffu, gh, m);
            gw2 += __shfl_xor_sync(0xffffffffu, gw2, m);
          }
          if (live && j < 3) {
            float* dop = p.doff + static_cast<size_t>(vox) * p.off_cstride + tap * 3;
            dop[j] = j == 0 ? gd : (j == 1 ? gh : gw2);
          }
        }
        __syncwarp();                                          // the tile is rewritten by the next tap
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kDMma) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 256); }
}

// ============================================================================================================
// weight gradient
// ============================================================================================================
constexpr int kWProd = 16;                               // gather warps
constexpr int kWThreads = (4 + 1 + kWProd) * 32;
constexpr int kWStages = 3;
constexpr int kWSmem = kWStages * kTile + 2 * kTile + (2 * kWStages + 2 * 2 + 1) * 8 + 16 + 128;
constexpr int kVoxPerPass = kWProd * 32 / 4;
constexpr int kMaxTapsPerLaunch = 14;

__global__ void __launch_bounds__(kWThreads, 1) dcn3d_bwd_weight_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  uint8_t* s_col = smem;                                  // [kWStages][kTile]  sampled A tiles
  uint8_t* s_dy = smem + kWStages * kTile;                // [2][kTile]
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_dy + 2 * kTile);
  uint64_t* bar_empty = bar_full + kWStages;
  uint64_t* bar_zfull = bar_empty + kWStages;
  uint64_t* bar_zempty = bar_zfull + 2;
  uint64_t* bar_done = bar_zempty + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWStages; ++i) { mbar_init(&bar_full[i], kWProd); mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_zfull[i], kWProd); mbar_init(&bar_zempty[i], 1); }
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 4) { tmem_alloc(s_tmem, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  if (warp < 4) {
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c0 = 0; c0 < 512; c0 += 16) tmem_zero16(lane_base + c0);
    tmem_st_wait();
    tc_fence_before_sync();
  }
  __syncthreads();
  tc_fence_after_sync();
  const int D = p.D, H = p.H, W = p.W, HW = p.H * p.W;
  const int per_cta = (p.nunits + gridDim.x - 1) / gridDim.x;
  const int unit_lo = min(static_cast<int>(blockIdx.x) * per_cta, p.nunits), unit_hi = min(unit_lo + per_cta, p.nunits);
  const int ntap = p.tap1 - p.tap0;

  if (warp > 4) {
    // ---------------- producers: dy tile per unit + the forward's trilinear gather per tap ------------------
    const int ptid = threadIdx.x - 5 * 32;
    const int cq = ptid & 3, vsub = ptid >> 2;
    constexpr int PASSES = 256 / kVoxPerPass;
    const uint32_t cs2 = static_cast<uint32_t>(p.x_cstride) * 2u;
    const char* xbytes = reinterpret_cast<const char*>(p.x);
    const uint32_t lane_off = cq * 32u;
    uint32_t g = 0, it = 0;
    for (int unit = unit_lo; unit < unit_hi; ++unit, ++it) {
      int ud, uh0, uw0, ub;
      unit_coords(unit, p, ud, uh0, uw0, ub);
      mbar_wait(&bar_zempty[it & 1], ((it >> 1) & 1u) ^ 1u);
      load_tile_256x64(s_dy + (it & 1) * kTile, p.dy, p, ud, uh0, uw0, ub, ptid, kWProd * 32);
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_zfull[it & 1]);
      int vh[PASSES], vw[PASSES];
      bool vlive[PASSES];
      const int vbase = ub * D * HW;
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int r = ps * kVoxPerPass + vsub;
        const int hh = uh0 + (r >> 4), ww = uw0 + (r & 15);
        vlive[ps] = (hh < H) && (ww < W);
        vh[ps] = vlive[ps] ? hh : 0;
        vw[ps] = vlive[ps] ? ww : 0;
      }
      for (int tap = p.tap0; tap < p.tap1; ++tap, ++g) {
        const int stage = g % kWStages;
        mbar_wait(&bar_empty[stage], ((g / kWStages) & 1u) ^ 1u);
        uint8_t* sa = s_col + stage * kTile;
        const int ti = tap / 9 - 1, tj = (tap / 3) % 3 - 1, tk = tap % 3 - 1;
        const float fdz = static_cast<float>(ud + ti);
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
          const int r = ps * kVoxPerPass + vsub;
          const int vox = vbase + (ud * H + vh[ps]) * W + vw[ps];
          const float* op = p.offset + static_cast<size_t>(vox) * p.off_cstride + tap * 3;
          const float pd = fdz + __ldg(op + 0);
          const float phh = static_cast<float>(vh[ps] + tj) + __ldg(op + 1);
          const float pw = static_cast<float>(vw[ps] + tk) + __ldg(op + 2);
          const bool inside = vlive[ps] && pd > -1.f && phh > -1.f && pw > -1.f && pd < static_cast<float>(D) &&
                              phh < static_cast<float>(H) && pw < static_cast<float>(W);
          const float fd = floorf(pd), fh = floorf(phh), fw = floorf(pw);
          const int d0 = static_cast<int>(fd), h0 = static_cast<int>(fh), w0 = static_cast<int>(fw);
          const float ld = pd - fd, lh = phh - fh, lw = pw - fw;
          const float wd0 = (inside && d0 >= 0) ? 1.f - ld : 0.f, wd1 = (inside && d0 + 1 <= D - 1) ? ld : 0.f;
          const float wh0 = (h0 >= 0) ? 1.f - lh : 0.f, wh1 = (h0 + 1 <= H - 1) ? lh : 0.f;
          const float ww0 = (w0 >= 0) ? 1.f - lw : 0.f, ww1 = (w0 + 1 <= W - 1) ? lw : 0.f;
          const int dc0 = min(max(d0, 0), D - 1), dc1 = min(max(d0 + 1, 0), D - 1);
          const int hc0 = min(max(h0, 0), H - 1), hc1 = min(max(h0 + 1, 0), H - 1);
          const int wc0 = min(max(w0, 0), W - 1), wc1 = min(max(w0 + 1, 0), W - 1);
          const uint32_t b00 = static_cast<uint32_t>(vbase + dc0 * HW + hc0 * W) * cs2 + lane_off;
          const uint32_t b01 = static_cast<uint32_t>(vbase + dc0 * HW + hc1 * W) * cs2 + lane_off;
          const uint32_t b10 = static_cast<uint32_t>(vbase + dc1 * HW + hc0 * W) * cs2 + lane_off;
          const uint32_t b11 = static_cast<uint32_t>(vbase + dc1 * HW + hc1 * W) * cs2 + lane_off;
          const uint32_t o0 = static_cast<uint32_t>(wc0) * cs2, o1 = static_cast<uint32_t>(wc1) * cs2;
          uint4 ua[8], ub2[8];
          ld_global_v8(xbytes + (b00 + o0), ua[0], ub2[0]);
          ld_global_v8(xbytes + (b00 + o1), ua[1], ub2[1]);
          ld_global_v8(xbytes + (b01 + o0), ua[2], ub2[2]);
          ld_global_v8(xbytes + (b01 + o1), ua[3], ub2[3]);
          ld_global_v8(xbytes + (b10 + o0), ua[4], ub2[4]);
          ld_global_v8(xbytes + (b10 + o1), ua[5], ub2[5]);
          ld_global_v8(xbytes + (b11 + o0), ua[6], ub2[6]);
          ld_global_v8(xbytes + (b11 + o1), ua[7], ub2[7]);
          const float a00 = wd0 * wh0, a01 = wd0 * wh1, a10 = wd1 * wh0, a11 = wd1 * wh1;
          const float cw[8] = {a00 * ww0, a00 * ww1, a01 * ww0, a01 * ww1, a10 * ww0, a10 * ww1, a11 * ww0, a11 * ww1};
          __nv_bfloat162 acc[8];
          {
            const __nv_bfloat162 w2 = __float2bfloat162_rn(cw[0]);
            acc[0] = __hmul2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ua[0].x));
            acc[1] = __hmul2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ua[0].y));
            acc[2] = __hmul2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ua[0].z));
            acc[3] = __hmul2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ua[0].w));
            acc[4] = __hmul2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ub2[0].x));
            acc[5] = __hmul2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ub2[0].y));
            acc[6] = __hmul2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ub2[0].z));
            acc[7] = __hmul2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ub2[0].w));
          }
#pragma unroll
          for (int c = 1; c < 8; ++c) {
            const __nv_bfloat162 w2 = __float2bfloat162_rn(cw[c]);
            acc[0] = __hfma2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ua[c].x), acc[0]);
            acc[1] = __hfma2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ua[c].y), acc[1]);
            acc[2] = __hfma2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ua[c].z), acc[2]);
            acc[3] = __hfma2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ua[c].w), acc[3]);
            acc[4] = __hfma2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ub2[c].x), acc[4]);
            acc[5] = __hfma2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ub2[c].y), acc[5]);
            acc[6] = __hfma2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ub2[c].z), acc[6]);
            acc[7] = __hfma2(w2, *reinterpret_cast<const __nv_bfloat162*>(&ub2[c].w), acc[7]);
          }
          uint4 oa, ob;
          oa.x = *reinterpret_cast<uint32_t*>(&acc[0]); oa.y = *reinterpret_cast<uint32_t*>(&acc[1]);
          oa.z = *reinterpret_cast<uint32_t*>(&acc[2]); oa.w = *reinterpret_cast<uint32_t*>(&acc[3]);
          ob.x = *reinterpret_cast<uint32_t*>(&acc[4]); ob.y = *reinterpret_cast<uint32_t*>(&acc[5]);
          ob.z = *reinterpret_cast<uint32_t*>(&acc[6]); ob.w = *reinterpret_cast<uint32_t*>(&acc[7]);
          uint8_t* dst = sa + (r >> 7) * kBlock + (2 * cq) * kChunk + (r & 127) * 16;
          *reinterpret_cast<uint4*>(dst) = oa;
          *reinterpret_cast<uint4*>(dst + kChunk) = ob;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[stage]);
      }
    }
  } else if (warp == 4) {
    // ---------------- MMA: dW_tap[c x o] += col_tap^T[c x 256 vox] * dy[256 vox x o]  (MN-major operands) ----
    const uint32_t idesc = umma_idesc_bf16_f32(64, kC) | (1u << 15) | (1u << 16);
    const uint64_t desc_hi = umma_desc_nosw(0, 128, kChunk);           // LBO = next 8 voxels, SBO = next channel chunk
    const bool leader = elect_one();
    uint32_t g = 0, it = 0;
    for (int unit = unit_lo; unit < unit_hi; ++unit, ++it) {
      mbar_wait(&bar_zfull[it & 1], (it >> 1) & 1u);
      tc_fence_after_sync();
      const uint32_t zt = smem_u32(s_dy + (it & 1) * kTile) >> 4;
      for (int j = 0; j < ntap; ++j, ++g) {
        const int stage = g % kWStages;
        mbar_wait(&bar_full[stage], (g / kWStages) & 1u);
        tc_fence_after_sync();
        if (leader) {
          const uint32_t at = smem_u32(s_col + stage * kTile) >> 4;
          const uint32_t acc = tmem_base + (static_cast<uint32_t>((j & 1) * 16) << 16) + (j >> 1) * kC;
#pragma unroll 4
          for (int kk = 0; kk < 16; ++kk) {
            const uint32_t o = (kk >> 3) * (kBlock >> 4) + (kk & 7) * 16;      // 16 rows per MMA, 16 B each
            umma_bf16(acc, desc_hi | static_cast<uint64_t>((at + o) & 0x3FFF), desc_hi | static_cast<uint64_t>((zt + o) & 0x3FFF), idesc, true);
          }
          umma_commit(&bar_empty[stage]);
          if (j == ntap - 1) umma_commit(&bar_zempty[it & 1]);
        }
        __syncwarp();
      }
    }
    if (leader) umma_commit(bar_done);
    __syncwarp();
  } else {
    mbar_wait(bar_done, 0);
    tc_fence_after_sync();
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int c = warp * 16 + (lane & 15), todd = lane >> 4;
    for (int jj = 0; jj < (ntap + 1) / 2; ++jj) {
      const int j = 2 * jj + todd;
#pragma unroll
      for (int c0 = 0; c0 < kC; c0 += 16) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(lane_base + jj * kC + c0, v);
        tmem_ld_wait();
        if (j < ntap) {
#pragma unroll
          for (int k = 0; k < 16; ++k) atomicAdd(p.dw + (static_cast<size_t>(p.tap0 + j) * kC + c) * kC + c0 + k, __uint_as_float(v[k]));
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

int fill(BwdParams& p, int B, int D, int H, int W, int x_cstride, int off_cstride) {
  p.B = B; p.D = D; p.H = H; p.W = W; p.x_cstride = x_cstride; p.off_cstride = off_cstride;
  p.tiles_h = (H + 15) / 16;
  p.tiles_w = (W + 15) / 16;
  p.nunits = B * D * p.tiles_h * p.tiles_w;
  return 0;
}

}  // namespace

extern "C" int dpf_dcn3d_bwd_data(const void* x, const float* offset, const void* dy, const void* w_t, float* dx, float* doffset,
                                  int B, int D, int H, int W, int x_cstride, int off_cstride, int dx_channels, void* stream) {
  DPF_REQUIRE(x && offset && dy && w_t && dx && doffset, "dpf_dcn3d_bwd_data: null pointer");
  DPF_REQUIRE(x_cstride >= kC && x_cstride % 8 == 0, "dpf_dcn3d_bwd_data: x_cstride=%d must be a multiple of 8 >= 64", x_cstride);
  DPF_REQUIRE(dx_channels == 32 || dx_channels == 64, "dpf_dcn3d_bwd_data: dx_channels=%d must be 32 or 64", dx_channels);
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(dx) && DPF_ALIGNED16(dy) && DPF_ALIGNED16(w_t), "dpf_dcn3d_bwd_data: pointers must be 16-byte aligned");
  DPF_REQUIRE(static_cast<long long>(B) * D * H * W < (1LL << 31) / 128, "dpf_dcn3d_bwd_data: tensor too large for 32-bit voxel indexing");
  BwdParams p{};
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.offset = offset;
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy);
  p.w = reinterpret_cast<const __nv_bfloat16*>(w_t);
  p.dx = dx; p.doff = doffset; p.dx_channels = dx_channels;
  DPF_REQUIRE(off_cstride >= 81, "dpf_dcn3d_bwd: off_cstride=%d must be >= 81", off_cstride);
  fill(p, B, D, H, W, x_cstride, off_cstride);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(dcn3d_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDSmem);
    if (e != cudaSuccess) return dpf::fail("dpf_dcn3d_bwd_data: shared memory opt-in: %s", cudaGetErrorString(e));
    attr = true;
  }
  dcn3d_bwd_data_kernel<<<std::min(p.nunits, dpf::sm_count()), kDThreads, kDSmem, static_cast<cudaStream_t>(stream)>>>(p);
  return dpf::after_launch("dpf_dcn3d_bwd_data");
}

extern "C" int dpf_dcn3d_bwd_weight(const void* x, const float* offset, const void* dy, float* dw, int B, int D, int H, int W,
                                    int x_cstride, int off_cstride, void* stream) {
  DPF_REQUIRE(x && offset && dy && dw, "dpf_dcn3d_bwd_weight: null pointer");
  DPF_REQUIRE(x_cstride >= kC && x_cstride % 8 == 0, "dpf_dcn3d_bwd_weight: x_cstride=%d must be a multiple of 8 >= 64", x_cstride);
  DPF_REQUIRE(static_cast<long long>(B) * D * H * W < (1LL << 31) / 128, "dpf_dcn3d_bwd_weight: tensor too large for 32-bit voxel indexing");
  BwdParams p{};
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.offset = offset;
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy);
  p.dw = dw;
  DPF_REQUIRE(off_cstride >= 81, "dpf_dcn3d_bwd: off_cstride=%d must be >= 81", off_cstride);
  fill(p, B, D, H, W, x_cstride, off_cstride);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(dcn3d_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem);
    if (e != cudaSuccess) return dpf::fail("dpf_dcn3d_bwd_weight: shared memory opt-in: %s", cudaGetErrorString(e));
    attr = true;
  }
  for (int t0 = 0; t0 < kTaps; t0 += kMaxTapsPerLaunch) {
    p.tap0 = t0;
    p.tap1 = std::min(kTaps, t0 + kMaxTapsPerLaunch);
    dcn3d_bwd_weight_kernel<<<std::min(p.nunits, dpf::sm_count()), kWThreads, kWSmem, static_cast<cudaStream_t>(stream)>>>(p);
    if (int rc = dpf::after_launch("dpf_dcn3d_bwd_weight")) return rc;
  }
  return 0;
}
