// 2-D 3x3 convolution (stride 1, any dilation d, padding d) on channels-last bf16 images as an implicit GEMM on tcgen05 tensor
// cores (sm_100a): Cin in {32, 64, 96}, Cout up to 96 in ONE launch, bias / scale / residual / LeakyReLU fused.
//
// Replaces the six `convtext` layers of the ANM normal branch (Conv2d 3x3, dilation 1,2,4,8,1,1, bias-free, + LeakyReLU(0.1):
// src/model/stereodpnet/normal_module.py:14-19,59-66 of the reference), which ran on cuDNN in round 1.
//
// Dilation by residue classes.  A 3x3 conv with dilation d only ever combines pixels whose row AND column indices are congruent
// mod d: the image splits into d*d independent sub-images (rows rh, rh+d, ...; columns rw, rw+d, ...), each of which sees a plain
// dilation-1 conv.  With channels-last storage a pixel is a contiguous 64-192 byte vector, so reading a sub-image (pixel stride
// d) costs nothing in coalescing.  A work tile is 16 x WT positions of ONE sub-image; the producer stages its 18 x (WT+2) halo
// window with 16-byte cp.async (zero-fill outside the sub-image = the conv padding), so any dilation runs at dilation-1 cost
// -- no window growth, no template parameter, one launch.
//
// GEMM view (same trick as conv3d_tc.cu): a GEMM block is 16 rows x 8 columns of the tile = 128 rows; the window is stored
// channel-chunk-planar, slot[c8][row][col][8 ch], which IS the UMMA no-swizzle K-major canonical layout (8 consecutive columns
// contiguous, row pitch = SBO, chunk-plane pitch = LBO): every tap (kh,kw) is a different descriptor START ADDRESS into the
// same window -- no im2col.  D[128 x Npad] += A[128 x 16] * W_tap[16 x Npad] over 9 taps x Cin/16 k-steps, fp32 in TMEM.
// Wide layers: N = Npad = Cout up to 96 in one MMA (A is read once per k-step for all output channels), and input channels are
// consumed in KPART-channel windows (slots of 32 or 48 channels) that accumulate into the same TMEM tile, so that the resident
// weights (up to 166 KB for 96 -> 96) and a 3-slot ring fit shared memory together.  All MMAs of a tile are issued by ONE
// thread in a fixed order: results are run-to-run deterministic.
//
// Pipeline: 4 producer warps (cp.async ring, full/empty mbarriers) -> 1 MMA warp (one elected lane) -> 4 epilogue warps
// (tcgen05.ld, scale/shift, residual, LeakyReLU, bf16 pack, 256-bit stores); accumulators double-buffered in TMEM; persistent
// CTAs, one per SM; the weights of all taps stay resident in shared memory.
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"
#include <algorithm>

namespace {

using namespace dpf;

constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kEpiWarps;                  // warp 4
constexpr int kProdWarps = 4;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;   // 288

struct Conv2dParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  int N, H, W, dil;
  int x_cstride, x_coff, y_cstride, y_coff;
  int cout;                      // channels stored (multiple of 8, <= NPAD)
  int cout_real;                 // channels of the layer (scale / shift have this many entries)
  int relu;
  float slope;
  int tiles_h, tiles_w, ntiles;  // per residue class: tiles over ceil(H/dil) x ceil(W/dil)
};

template <int CIN, int KPART, int NPAD, int WT, int NS>
struct C2Cfg {
  static constexpr int NCHP = KPART / 8;                       // 16-byte channel chunks per slot
  static constexpr int KPARTS = CIN / KPART;
  static constexpr int KSTEPS = KPART / 16;
  static constexpr int NBLK = WT / 8;
  static constexpr int WP = WT + 2;
  static constexpr int PLANE_BYTES = 18 * WP * 16;
  // chunk-plane pitch: 16-byte multiple whose residue mod 128 spreads a quarter-warp's cp.async writes over the banks
  static constexpr int WANT = (NCHP == 4) ? 32 : (NCHP == 6 ? 112 : 16);
  static constexpr int CH_STRIDE = PLANE_BYTES + ((WANT - (PLANE_BYTES % 128)) + 128) % 128;
  static constexpr int SLOT_BYTES = NCHP * CH_STRIDE;
  static constexpr int W_TAP_BYTES = (CIN / 8) * NPAD * 16;    // [c8][NPAD][8]
  static constexpr int W_BYTES = 9 * W_TAP_BYTES;
  static constexpr int ACC_COLS = NBLK * NPAD;
  static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = W_BYTES + NS * SLOT_BYTES + 2 * NPAD * 4 + (2 * NS + 4) * 8 + 16 + 128;
  static_assert(CIN % KPART == 0 && KPART % 16 == 0, "input channels are consumed in KPART windows of whole k-steps");
  static_assert(NPAD % 16 == 0 && NPAD >= 16 && NPAD <= 256, "UMMA N (M = 128): multiple of 16 in [16, 256]");
  static_assert(2 * ACC_COLS <= 512, "accumulators do not fit TMEM");
  static_assert(CH_STRIDE % 16 == 0, "chunk pitch must be a 16-byte multiple");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

__device__ __forceinline__ void decode_tile(int tile, const Conv2dParams& p, int& n, int& rh, int& rw, int& th, int& tw) {
  rw = tile % p.dil;
  int t = tile / p.dil;
  rh = t % p.dil;
  t /= p.dil;
  tw = t % p.tiles_w;
  t /= p.tiles_w;
  th = t % p.tiles_h;
  n = t / p.tiles_h;
}

template <int CIN, int KPART, int NPAD, int WT, int NS>
__global__ void __launch_bounds__(kThreads, 1) conv2d_tc_kernel(const __grid_constant__ Conv2dParams p) {
  using C = C2Cfg<CIN, KPART, NPAD, WT, NS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  uint8_t* s_w = smem;
  uint8_t* s_slots = smem + C::W_BYTES;
  float* s_scale = reinterpret_cast<float*>(s_slots + NS * C::SLOT_BYTES);
  float* s_shift = s_scale + NPAD;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_shift + NPAD);
  uint64_t* bar_empty = bar_full + NS;
  uint64_t* bar_tfull = bar_empty + NS;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup: resident weights [tap][c8][NPAD][8], epilogue affine, barriers, TMEM ------------------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = threadIdx.x; i < C::W_BYTES / 16; i += kThreads) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < NPAD; i += kThreads) {
      s_scale[i] = (p.scale != nullptr && i < p.cout_real) ? p.scale[i] : 1.0f;
      s_shift[i] = (p.shift != nullptr && i < p.cout_real) ? p.shift[i] : 0.0f;
    }
    fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&bar_full[i], kProdWarps);
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
  const int H = p.H, W = p.W, d = p.dil;

  if (warp > kMmaWarp) {
    // =================================== producers: global -> shared ring =========================================
    const int pwarp = warp - (kMmaWarp + 1);
    constexpr int PIECES_PER_ROW = C::WP * C::NCHP;
    uint32_t g = 0;
    int prev_slot = -1;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int n, rh, rw, th, tw;
      decode_tile(tile, p, n, rh, rw, th, tw);
      const int Hs = (H - rh + d - 1) / d, Ws = (W - rw + d - 1) / d;      // extent of this residue class's sub-image
      const int hs0 = th * 16 - 1, ws0 = tw * WT - 1;                      // window origin in sub-image coordinates
      const __nv_bfloat16* ximg = p.x + static_cast<size_t>(n) * H * W * p.x_cstride + p.x_coff;
      for (int kp = 0; kp < C::KPARTS; ++kp, ++g) {
        const int slot = g % NS;
        mbar_wait(&bar_empty[slot], ((g / NS) & 1u) ^ 1u);
        const uint32_t sbase = smem_u32(s_slots + slot * C::SLOT_BYTES);
        for (int row = pwarp; row < 18; row += kProdWarps) {
          const int hs = hs0 + row;
          const bool hok = (hs >= 0) && (hs < Hs);
          const __nv_bfloat16* xrow = ximg + (hok ? static_cast<size_t>(hs * d + rh) * W * p.x_cstride : 0) + kp * KPART;
#pragma unroll
          for (int q = lane; q < PIECES_PER_ROW; q += 32) {
            const int col = q / C::NCHP, c8 = q - col * C::NCHP;
            const int ws = ws0 + col;
            const bool ok = hok && (ws >= 0) && (ws < Ws);
            const __nv_bfloat16* src = ok ? (xrow + static_cast<size_t>(ws * d + rw) * p.x_cstride + c8 * 8) : p.x;
            cp_async16_zfill(sbase + c8 * C::CH_STRIDE + (row * C::WP + col) * 16, src, ok);
          }
        }
        cp_async_commit();
        if (prev_slot >= 0) {                                    // complete the previous window (one group of lag)
          cp_async_wait<1>();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_full[prev_slot]);
        }
        prev_slot = slot;
      }
    }
    if (prev_slot >= 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[prev_slot]);
    }
  } else if (warp == kMmaWarp) {
    // =================================== MMA issuer (one elected lane, fixed order) ===============================
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, NPAD);
    const uint32_t wbase = smem_u32(s_w) >> 4;
    const uint32_t sbase0 = smem_u32(s_slots);
    const uint64_t adesc_hi = umma_desc_nosw(0, C::CH_STRIDE, C::WP * 16);       // LBO = chunk-plane pitch, SBO = window row pitch
    const uint64_t bdesc_hi = umma_desc_nosw(0, NPAD * 16, 128);
    const bool leader = elect_one();
    uint32_t g = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      int n, rh, rw, th, tw;
      decode_tile(tile, p, n, rh, rw, th, tw);
      const int Ws = (W - rw + d - 1) / d;
      const int nblk = max(0, min(C::NBLK, (Ws - tw * WT + 7) >> 3));
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      mbar_wait(&bar_tempty[as], aph ^ 1u);
      tc_fence_after_sync();
      for (int kp = 0; kp < C::KPARTS; ++kp, ++g) {
        const uint32_t slot = g % NS;
        mbar_wait(&bar_full[slot], (g / NS) & 1u);
        tc_fence_after_sync();
        if (leader) {
          const uint32_t a_slot = (sbase0 + slot * C::SLOT_BYTES) >> 4;
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const uint32_t a0 = a_slot + kh * C::WP + kw;
              const uint32_t b0 = wbase + (kh * 3 + kw) * (C::W_TAP_BYTES >> 4) + kp * C::NCHP * NPAD;
#pragma unroll
              for (int ks = 0; ks < C::KSTEPS; ++ks) {
                const uint64_t bdesc = bdesc_hi | static_cast<uint64_t>((b0 + ks * 2 * NPAD) & 0x3FFF);
#pragma unroll
                for (int blk = 0; blk < C::NBLK; ++blk) {
                  if (blk < nblk) {                           // (nblk can be 0 here: tiles beyond a short residue class)
                    const uint64_t adesc = adesc_hi | static_cast<uint64_t>((a0 + ks * 2 * (C::CH_STRIDE >> 4) + blk * 8) & 0x3FFF);
                    umma_bf16(tmem_base + (as * C::NBLK + blk) * NPAD, adesc, bdesc, idesc, (kp | kh | kw | ks) != 0);
                  }
                }
              }
            }
          }
          umma_commit(&bar_empty[slot]);
          if (kp == C::KPARTS - 1) umma_commit(&bar_tfull[as]);
        }
        __syncwarp();
      }
    }
  } else {
    // =================================== epilogue: TMEM -> registers -> global ====================================
    uint32_t it = 0;
    const int m = warp * 32 + lane;
    const int hrow = m >> 3, wcol = m & 7;
    const bool wide = ((p.y_cstride | p.y_coff) & 15) == 0;                   // 32-byte aligned 16-channel chunks
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      int n, rh, rw, th, tw;
      decode_tile(tile, p, n, rh, rw, th, tw);
      const int Hs = (H - rh + d - 1) / d, Ws = (W - rw + d - 1) / d;
      const int nblk = max(0, min(C::NBLK, (Ws - tw * WT + 7) >> 3));
      const int hs = th * 16 + hrow;
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      // narrow layers (N <= 32): the residual of the tile is requested before the accumulator wait (it does not depend on the MMAs;
      // fetched inside the store loop it is one dependent global round trip per block on the four epilogue warps)
      constexpr bool kPref = (NPAD <= 32);
      constexpr int PCH = kPref ? NPAD / 16 : 1;
      uint4 rr[kPref ? C::NBLK : 1][PCH][2];
      const bool res_pref = kPref && (p.residual != nullptr) && wide && (p.cout % 16 == 0);
      if (kPref && res_pref) {
#pragma unroll
        for (int blk = 0; blk < C::NBLK; ++blk) {
          const int ws = tw * WT + blk * 8 + wcol;
          const bool ok = (blk < nblk) && (hs < Hs) && (ws < Ws);
          const size_t pix = (static_cast<size_t>(n) * H + (ok ? hs * d + rh : 0)) * W + (ok ? ws * d + rw : 0);
#pragma unroll
          for (int q = 0; q < PCH; ++q) {
            rr[kPref ? blk : 0][q][0] = make_uint4(0u, 0u, 0u, 0u);
            rr[kPref ? blk : 0][q][1] = rr[kPref ? blk : 0][q][0];
            if (ok && q * 16 < p.cout)
              ld_global_v8(p.residual + pix * p.y_cstride + p.y_coff + q * 16, rr[kPref ? blk : 0][q][0], rr[kPref ? blk : 0][q][1]);
          }
        }
      }
      mbar_wait(&bar_tfull[as], aph);
      tc_fence_after_sync();
#pragma unroll(kPref ? C::NBLK : 1)
      for (int blk = 0; blk < (kPref ? C::NBLK : nblk); ++blk) {
        if (kPref && blk >= nblk) continue;
        const int ws = tw * WT + blk * 8 + wcol;
        const bool ok = (hs < Hs) && (ws < Ws);
        const size_t pix = (static_cast<size_t>(n) * H + (ok ? hs * d + rh : 0)) * W + (ok ? ws * d + rw : 0);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + (as * C::NBLK + blk) * NPAD;
#pragma unroll
        for (int c0 = 0; c0 < NPAD; c0 += 16) {
          uint32_t v[16];
          __syncwarp();                                        // tcgen05.ld is .sync.aligned: keep the warp converged
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (!ok || c0 >= p.cout) continue;
          const int nst = min(16, p.cout - c0);                // 8 or 16 channels stored
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * s_scale[c0 + j] + s_shift[c0 + j];
          __nv_bfloat16* yo = p.y + pix * p.y_cstride + p.y_coff + c0;
          const bool res_post = (p.relu & 2) != 0;             // y = act(v) + residual instead of act(v + residual)
          if (res_post && (p.relu & 1)) {
            const float sl = p.slope;
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f) + sl * fminf(f[j], 0.f);
          }
          if (p.residual != nullptr) {
            const __nv_bfloat16* ro = p.residual + pix * p.y_cstride + p.y_coff + c0;
            uint4 r0, r1 = make_uint4(0u, 0u, 0u, 0u);
            if (kPref && res_pref) {
              r0 = rr[kPref ? blk : 0][kPref ? (c0 >> 4) : 0][0];
              r1 = rr[kPref ? blk : 0][kPref ? (c0 >> 4) : 0][1];
            } else {
              r0 = *reinterpret_cast<const uint4*>(ro);
              if (nst > 8) r1 = *reinterpret_cast<const uint4*>(ro + 8);
            }
            f[0] += bf16_lo(r0.x); f[1] += bf16_hi(r0.x); f[2] += bf16_lo(r0.y); f[3] += bf16_hi(r0.y);
            f[4] += bf16_lo(r0.z); f[5] += bf16_hi(r0.z); f[6] += bf16_lo(r0.w); f[7] += bf16_hi(r0.w);
            f[8] += bf16_lo(r1.x); f[9] += bf16_hi(r1.x); f[10] += bf16_lo(r1.y); f[11] += bf16_hi(r1.y);
            f[12] += bf16_lo(r1.z); f[13] += bf16_hi(r1.z); f[14] += bf16_lo(r1.w); f[15] += bf16_hi(r1.w);
          }
          if ((p.relu & 1) && !res_post) {
            const float sl = p.slope;                           // 0 = ReLU, else LeakyReLU / single-parameter PReLU
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f) + sl * fminf(f[j], 0.f);
          }
          uint4 o0, o1;
          o0.x = pack_bf16x2(f[0], f[1]); o0.y = pack_bf16x2(f[2], f[3]); o0.z = pack_bf16x2(f[4], f[5]); o0.w = pack_bf16x2(f[6], f[7]);
          o1.x = pack_bf16x2(f[8], f[9]); o1.y = pack_bf16x2(f[10], f[11]); o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
          if (wide && nst == 16) st_global_v8(yo, o0, o1);
          else {
            *reinterpret_cast<uint4*>(yo) = o0;
            if (nst > 8) *reinterpret_cast<uint4*>(yo + 8) = o1;
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[as]);
    }
  }

  // ---- teardown ----------------------------------------------------------------------------------------------------
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int CIN, int KPART, int NPAD, int WT, int NS>
int launch2d(Conv2dParams kp, cudaStream_t st) {
  using C = C2Cfg<CIN, KPART, NPAD, WT, NS>;
  const int hs = (kp.H + kp.dil - 1) / kp.dil, ws = (kp.W + kp.dil - 1) / kp.dil;
  kp.tiles_h = (hs + 15) / 16;
  kp.tiles_w = (ws + WT - 1) / WT;
  const long long nt = static_cast<long long>(kp.N) * kp.dil * kp.dil * kp.tiles_h * kp.tiles_w;
  if (nt >= (1LL << 31)) return dpf::fail("dpf_conv2d_tc_fwd: too many tiles");
  kp.ntiles = static_cast<int>(nt);
  auto kern = conv2d_tc_kernel<CIN, KPART, NPAD, WT, NS>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return dpf::fail("dpf_conv2d_tc_fwd: cannot opt in to %d B shared memory: %s", C::SMEM_BYTES, cudaGetErrorString(e));
    attr_done = true;
  }
  const int grid = std::min(kp.ntiles, dpf::sm_count());
  kern<<<grid, kThreads, C::SMEM_BYTES, st>>>(kp);
  return dpf::after_launch("dpf_conv2d_tc_fwd");
}

}  // namespace

extern "C" int dpf_conv2d_tc_npad(int Cout) { return (Cout + 15) / 16 * 16; }

extern "C" long long dpf_conv2d_tc_weight_elems(int Cin, int Cout) {
  return 9LL * Cin * dpf_conv2d_tc_npad(Cout);
}

extern "C" int dpf_conv2d_tc_fwd(const void* x, const void* w, void* y, const float* scale, const float* shift, const void* residual,
                                 int N, int H, int W, int Cin, int Cout, int x_cstride, int x_coff, int y_cstride, int y_coff,
                                 int dil, int relu, float slope, void* stream) {
  DPF_REQUIRE(x && w && y, "dpf_conv2d_tc_fwd: null tensor pointer");
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(w) && DPF_ALIGNED16(y) && (residual == nullptr || DPF_ALIGNED16(residual)),
              "dpf_conv2d_tc_fwd: pointers must be 16-byte aligned");
  DPF_REQUIRE(N > 0 && H > 0 && W > 0 && dil >= 1 && dil <= 64, "dpf_conv2d_tc_fwd: bad shape / dilation");
  DPF_REQUIRE(Cout >= 1 && Cout <= 96, "dpf_conv2d_tc_fwd: Cout=%d must be in [1, 96]", Cout);
  DPF_REQUIRE(x_cstride % 8 == 0 && x_coff % 8 == 0 && x_coff + Cin <= x_cstride, "dpf_conv2d_tc_fwd: bad input channel window");
  const int cstore = (Cout + 7) / 8 * 8;       // channels written: Cout rounded up to a 16-byte piece (extra ones are exact zeros)
  DPF_REQUIRE(y_cstride % 8 == 0 && y_coff % 8 == 0 && y_coff + cstore <= y_cstride,
              "dpf_conv2d_tc_fwd: the output needs room for %d channels at y_coff (y_cstride %d)", cstore, y_cstride);
  DPF_REQUIRE(static_cast<long long>(N) * H * W < (1LL << 31), "dpf_conv2d_tc_fwd: tensor too large for 32-bit pixel indexing");
  Conv2dParams kp{};
  kp.x = reinterpret_cast<const __nv_bfloat16*>(x);
  kp.w = reinterpret_cast<const __nv_bfloat16*>(w);
  kp.y = reinterpret_cast<__nv_bfloat16*>(y);
  kp.scale = scale; kp.shift = shift;
  kp.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  kp.N = N; kp.H = H; kp.W = W; kp.dil = dil;
  kp.x_cstride = x_cstride; kp.x_coff = x_coff; kp.y_cstride = y_cstride; kp.y_coff = y_coff;
  kp.cout = cstore; kp.cout_real = Cout; kp.relu = relu; kp.slope = slope;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int npad = dpf_conv2d_tc_npad(Cout);
  // (Cin, Npad) -> <CIN, KPART, NPAD, WT, NS>: the widest tile whose weights + 3-4 slot ring fit the 227 KB of shared memory
  if (Cin == 32) {
    if (npad == 16) return launch2d<32, 32, 16, 24, 3>(kp, st);
    if (npad == 32) return launch2d<32, 32, 32, 24, 3>(kp, st);
    if (npad == 48) return launch2d<32, 32, 48, 24, 3>(kp, st);
    if (npad == 64) return launch2d<32, 32, 64, 24, 3>(kp, st);
    if (npad == 96) return launch2d<32, 32, 96, 16, 3>(kp, st);
  } else if (Cin == 64) {
    if (npad == 16) return launch2d<64, 32, 16, 16, 4>(kp, st);
    if (npad == 32) return launch2d<64, 32, 32, 16, 4>(kp, st);
    if (npad == 64) return launch2d<64, 32, 64, 16, 4>(kp, st);
    if (npad == 96) return launch2d<64, 32, 96, 16, 3>(kp, st);
  } else if (Cin == 96) {
    if (npad == 32) return launch2d<96, 48, 32, 16, 3>(kp, st);
    if (npad == 64) return launch2d<96, 48, 64, 16, 3>(kp, st);
    if (npad == 96) return launch2d<96, 48, 96, 8, 3>(kp, st);
  }
  return dpf::fail("dpf_conv2d_tc_fwd: no kernel for Cin=%d Cout=%d (built: Cin 32 | 64 | 96; Cout <= 16 | 32 | 48 | 64 | 96 classes)", Cin, Cout);
}
