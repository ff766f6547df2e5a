// 32 -> 1 channel 3x3x3 convolution (stride 1, pad 1) for the hourglass heads, as a bandwidth-bound tcgen05 kernel (sm_100a).
//
// Replaces the `nn.Conv3d(32, 1, 3, 1, 1, bias=False)` that closes every classifK of PSMNetHGAggregation
// (src/model/stereodpnet/modules.py:288-296; cumulative adds :323-325) and StereoNet's conv3d_alone
// (src/model/stereonet/mainmodel.py:50-51,119).  On the generic engine this layer is 27 taps x (128 x 16 x 32) MMAs whose cost is
// the 27 A-operand reads from shared memory (0.163 ms per launch at 4 x 8 x 280 x 420, 2.8 % of the tensor peak, round 1).
//
// Here the TAPS are the GEMM's N dimension:  P[pos, tap] = sum_c x[pos, c] * W[tap, c]  -- one pair of M128 x N32 x K16 MMAs per
// 128 positions, the activation tile is read from shared memory ONCE -- and the 27 partial planes are then combined with their
// spatial shifts on the CUDA cores:  y[d, h, w] = sum_{kd,kh,kw} P_{d+kd-1}[(h+kh-1, w+kw-1), (kd,kh,kw)].
// A CTA owns a 14 x 30 output tile (16 x 32 input region, 4 GEMM blocks of 16 rows x 8 columns) and streams the D input planes:
// plane z is staged once (cp.async, 32 KB), multiplied (8 MMAs), its P drained from TMEM to shared memory by four warps
// ([27][16 x 40] fp32, conflict-free pitch, double-buffered) and added by four other warps into the register accumulators of
// output planes z-1, z, z+1; output plane z-1 is then complete and stored (+ shift + residual, fp32).  HBM traffic = the input once (x 512/420 halo overhead) + 4 (+4) B per output voxel.
// One MMA-issuing thread and a fixed summation order: deterministic.
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"
#include <algorithm>

namespace {

using namespace dpf;

constexpr int kDrainWarps = 4, kSumWarps = 8;                    // TMEM -> P planes | shifted sums -> output
constexpr int kEpiWarps = kDrainWarps + kSumWarps;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kProdWarps = 4;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;     // 544
constexpr int RH = 16, RW = 32;                                  // input region of a tile (rows x columns)
constexpr int OH = RH - 2, OW = RW - 2;                          // output tile
constexpr int NBLK = RW / 8;                                     // GEMM blocks (16 rows x 8 columns = 128 positions)
constexpr int NS = 4;                                            // input slots
constexpr int NP = 1;                                            // P buffers
constexpr int CH_STRIDE = RH * RW * 16 + 32;                     // bytes between 8-channel chunk planes (+32: conflict-free cp.async)
constexpr int SLOT_BYTES = 4 * CH_STRIDE;
constexpr int W_BYTES = 4 * 32 * 16;                             // [c8][32 taps][8] bf16
constexpr int PP = 40;                                           // row pitch of a P plane in floats (rows 4q..4q+3 -> distinct banks)
constexpr int P_TAP = RH * PP;                                   // floats per tap plane
constexpr int P_BYTES = 27 * P_TAP * 4;
constexpr int SMEM_BYTES = W_BYTES + NS * SLOT_BYTES + NP * P_BYTES + (2 * NS + 8) * 8 + 16 + 128;
static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");

struct HeadParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
  float* y;
  const float* residual;
  float shift;
  int B, D, H, W, x_cstride;
  int tiles_h, tiles_w, ntiles;
};


__global__ void __launch_bounds__(kThreads, 1) conv3d_head_kernel(const __grid_constant__ HeadParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  uint8_t* s_w = smem;
  uint8_t* s_slots = smem + W_BYTES;
  float* s_p = reinterpret_cast<float*>(s_slots + NS * SLOT_BYTES);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_p) + NP * P_BYTES);
  uint64_t* bar_empty = bar_full + NS;
  uint64_t* bar_tfull = bar_empty + NS;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint64_t* bar_pfull = bar_tempty + 2;                           // P buffer written by the drain warps
  uint64_t* bar_pempty = bar_pfull + 2;                           // ... consumed by the sum warps
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_pempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = threadIdx.x; i < W_BYTES / 16; i += kThreads) dst[i] = __ldg(src + i);
    fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&bar_full[i], kProdWarps);
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], kDrainWarps);
      mbar_init(&bar_pfull[i], kDrainWarps);
      mbar_init(&bar_pempty[i], kSumWarps);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(s_tmem, 256);                                      // 2 stages x 4 blocks x 32 columns
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const int D = p.D, H = p.H, W = p.W;

  if (warp > kMmaWarp) {
    // =================================== producers: (tile, plane) -> slot ring ======================================
    const int ptid = threadIdx.x - (kMmaWarp + 1) * 32;          // 0..127
    uint32_t g = 0;
    auto land = [&](int slot) {                                   // this thread's copies into `slot` have landed: publish them
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[slot]);
    };
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int b = tile / (p.tiles_w * p.tiles_h);
      const int h0 = th * OH - 1, w0 = tw * OW - 1;
      for (int z = 0; z < D; ++z, ++g) {
        const int slot = g % NS;
        mbar_wait(&bar_empty[slot], ((g / NS) & 1u) ^ 1u);
        const uint32_t sbase = smem_u32(s_slots + slot * SLOT_BYTES);
        const __nv_bfloat16* xplane = p.x + (static_cast<size_t>(b) * D + z) * H * static_cast<size_t>(W) * p.x_cstride;
#pragma unroll 4
        for (int i = ptid; i < RH * RW * 4; i += kProdWarps * 32) {
          const int pos = i >> 2, c8 = i & 3;
          const int r = pos >> 5, c = pos & 31;
          const int h = h0 + r, w = w0 + c;
          const bool ok = (h >= 0) && (h < H) && (w >= 0) && (w < W);
          const __nv_bfloat16* src = ok ? xplane + (static_cast<size_t>(h) * W + w) * p.x_cstride + c8 * 8 : p.x;
          cp_async16_zfill(sbase + c8 * CH_STRIDE + pos * 16, src, ok);
        }
        cp_async_commit();
        if (g >= 2) {                                             // three planes in flight: plane g-2 is published now
          cp_async_wait<2>();
          land((g - 2) % NS);
        }
      }
    }
    if (g >= 2) {
      cp_async_wait<1>();
      land((g - 2) % NS);
    }
    if (g >= 1) {
      cp_async_wait<0>();
      land((g - 1) % NS);
    }
  } else if (warp == kMmaWarp) {
    // =================================== MMA issuer: 4 blocks x 2 k-steps per plane ================================
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, 32);
    const uint32_t wbase = smem_u32(s_w) >> 4;
    const uint32_t sbase0 = smem_u32(s_slots);
    const uint64_t adesc_hi = umma_desc_nosw(0, CH_STRIDE, RW * 16);    // LBO = chunk plane, SBO = region row
    const uint64_t bdesc_hi = umma_desc_nosw(0, 32 * 16, 128);
    const bool leader = elect_one();
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      for (int z = 0; z < D; ++z, ++g) {
        const uint32_t slot = g % NS, as = g & 1u;
        mbar_wait(&bar_tempty[as], ((g >> 1) & 1u) ^ 1u);
        mbar_wait(&bar_full[slot], (g / NS) & 1u);
        tc_fence_after_sync();
        if (leader) {
          const uint32_t a_slot = (sbase0 + slot * SLOT_BYTES) >> 4;
#pragma unroll
          for (int blk = 0; blk < NBLK; ++blk) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t adesc = adesc_hi | static_cast<uint64_t>((a_slot + blk * 8 + ks * 2 * (CH_STRIDE >> 4)) & 0x3FFF);
              const uint64_t bdesc = bdesc_hi | static_cast<uint64_t>((wbase + ks * 2 * 32) & 0x3FFF);
              umma_bf16(tmem_base + as * 128 + blk * 32, adesc, bdesc, idesc, ks != 0);
            }
          }
          umma_commit(&bar_empty[slot]);
          umma_commit(&bar_tfull[as]);
        }
        __syncwarp();
      }
    }
  } else if (warp < kDrainWarps) {
    // =================================== drain warps: TMEM accumulators -> P planes in shared memory =================
    // GEMM row m of block j = region row m / 8, column 8 j + m % 8; warp q owns TMEM lanes 32 q .. 32 q + 31 of all 4 blocks
    const int q = warp;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      for (int z = 0; z < D; ++z, ++g) {
        const uint32_t as = g & 1u;
        mbar_wait(&bar_tfull[as], (g >> 1) & 1u);
        const uint32_t ps = g % NP;
        mbar_wait(&bar_pempty[ps], ((g / NP) & 1u) ^ 1u);
        tc_fence_after_sync();
        float* pbuf = s_p + ps * (P_BYTES / 4);
        const int m = q * 32 + lane;
#pragma unroll
        for (int blk = 0; blk < NBLK; ++blk) {
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 128 + blk * 32;
          uint32_t v0[16], v1[16];
          tmem_ld16(taddr, v0);
          tmem_ld16(taddr + 16, v1);
          tmem_ld_wait();
          float* dst = pbuf + (m >> 3) * PP + blk * 8 + (m & 7);
#pragma unroll
          for (int t = 0; t < 16; ++t) dst[t * P_TAP] = __uint_as_float(v0[t]);
#pragma unroll
          for (int t = 0; t < 11; ++t) dst[(16 + t) * P_TAP] = __uint_as_float(v1[t]);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bar_tempty[as]);                           // accumulator stage free for the MMA of plane z + 2
          mbar_arrive(&bar_pfull[ps]);                            // (release: the st.shared above are ordered before the arrive)
        }
      }
    }
  } else {
    // =================================== sum warps: shifted sums over the 27 P planes -> output =======================
    constexpr int NR = 16 / kSumWarps;                            // output rows per thread
    const int et = threadIdx.x - kDrainWarps * 32;                // 0..255
    const int oc = et & 31;                                       // region column of this thread's outputs
    const int orow0 = 1 + (et >> 5);                              // region rows orow0 + kSumWarps * i, i = 0..NR-1
    const bool col_ok = (oc >= 1) && (oc <= OW);
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int b = tile / (p.tiles_w * p.tiles_h);
      const int h0 = th * OH - 1, w0 = tw * OW - 1;
      float acc[NR][3];                                           // [output row][plane z-1, z, z+1]
#pragma unroll
      for (int i = 0; i < NR; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;
      for (int z = 0; z < D; ++z, ++g) {
        const uint32_t as = g & 1u;
        // the residual of the output plane that completes in this iteration is requested before the wait (a global round trip)
        float rsd[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          const int orow = orow0 + kSumWarps * i;
          const int h = h0 + orow, w = w0 + oc;
          rsd[i] = 0.f;
          if (p.residual != nullptr && z >= 1 && col_ok && orow <= OH && h < H && w < W)
            rsd[i] = __ldg(p.residual + ((static_cast<size_t>(b) * D + (z - 1)) * H + h) * W + w);
        }
        const uint32_t ps = g % NP;
        mbar_wait(&bar_pfull[ps], (g / NP) & 1u);
        const float* pbuf = s_p + ps * (P_BYTES / 4);
        // ---- input plane z feeds output planes z+1 (kd 0), z (kd 1), z-1 (kd 2)
        if (col_ok) {
#pragma unroll
          for (int i = 0; i < NR; ++i) {
            const int orow = orow0 + kSumWarps * i;
            if (orow <= OH) {
              const float* base = pbuf + (orow - 1) * PP + (oc - 1);
#pragma unroll
              for (int kd = 0; kd < 3; ++kd) {
                float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                  s0 += base[((kd * 3 + kh) * 3 + 0) * P_TAP + kh * PP + 0];
                  s1 += base[((kd * 3 + kh) * 3 + 1) * P_TAP + kh * PP + 1];
                  s2 += base[((kd * 3 + kh) * 3 + 2) * P_TAP + kh * PP + 2];
                }
                acc[i][2 - kd] += (s0 + s1) + s2;
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_pempty[ps]);              // this warp has read everything it needs from the buffer
        // ---- output plane z-1 is complete (and plane D-1 after the last input plane)
#pragma unroll
        for (int fin = 0; fin < 2; ++fin) {
          const int zo = (fin == 0) ? z - 1 : z;
          if (fin == 1 && z != D - 1) break;
          if (zo >= 0 && col_ok) {
#pragma unroll
            for (int i = 0; i < NR; ++i) {
              const int orow = orow0 + kSumWarps * i;
              const int h = h0 + orow, w = w0 + oc;
              if (orow <= OH && h < H && w < W) {
                const size_t o = ((static_cast<size_t>(b) * D + zo) * H + h) * W + w;
                float val = acc[i][fin] + p.shift;
                if (fin == 0) val += rsd[i];
                else if (p.residual != nullptr) val += __ldg(p.residual + o);
                p.y[o] = val;
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          acc[i][0] = acc[i][1];
          acc[i][1] = acc[i][2];
          acc[i][2] = 0.f;
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

extern "C" int dpf_conv3d_head_fwd(const void* x, const void* w, float* y, const float* residual, float shift, int B, int D, int H,
                                   int W, int x_cstride, void* stream) {
  DPF_REQUIRE(x && w && y, "dpf_conv3d_head_fwd: null pointer");
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(w), "dpf_conv3d_head_fwd: x and w must be 16-byte aligned");
  DPF_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "dpf_conv3d_head_fwd: bad shape");
  DPF_REQUIRE(x_cstride >= 32 && x_cstride % 8 == 0, "dpf_conv3d_head_fwd: x_cstride=%d must be a multiple of 8 >= 32", x_cstride);
  HeadParams p{};
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.w = reinterpret_cast<const __nv_bfloat16*>(w);
  p.y = y; p.residual = residual; p.shift = shift;
  p.B = B; p.D = D; p.H = H; p.W = W; p.x_cstride = x_cstride;
  p.tiles_h = (H + OH - 1) / OH;
  p.tiles_w = (W + OW - 1) / OW;
  const long long nt = static_cast<long long>(B) * p.tiles_h * p.tiles_w;
  DPF_REQUIRE(nt < (1LL << 31), "dpf_conv3d_head_fwd: too many tiles");
  p.ntiles = static_cast<int>(nt);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv3d_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return dpf::fail("dpf_conv3d_head_fwd: cannot opt in to %d B shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
    attr_done = true;
  }
  const int grid = std::min(p.ntiles, dpf::sm_count());
  conv3d_head_kernel<<<grid, kThreads, SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(p);
  return dpf::after_launch("dpf_conv3d_head_fwd");
}
