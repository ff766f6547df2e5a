// Weight gradient of the stride-1 3-D convolutions on tcgen05 tensor cores (sm_100a).
//
// Replaces the dW half of autograd through nn.Conv3d in the reference's aggregation
// (src/model/stereodpnet/modules.py:204-337):   dW[tap][ci][co] = sum_{b,d,h,w} x[b, d+kd-1, h+kh-1, w+kw-1, ci] * dz[b,d,h,w,co].
//
// GEMM view per tap: D[ci x co] += X_tap^T [ci x P] * dZ [P x co] with the voxel positions P as the K dimension.  Both
// operands are read from the SAME channel-chunk-planar shared-memory staging as the forward kernel
// (slot[c8][row][col][8 channels]) -- here as MN-major UMMA operands: the channel index is the contiguous (M / N)
// direction, 8 consecutive w positions are the 8 K-rows of a core matrix (16 B apart), the next 8 positions are LBO = 128 B
// further, channel chunks are SBO = chunk-plane pitch apart.  One tcgen05.mma (M = 64 input channels, N = Cout, K = 16
// positions) therefore consumes 16 consecutive voxels of one row for one tap; the tap shift is, as in the forward
// kernel, only a different start address of the A operand.  27 accumulators [64 x Cout] fp32 live in TMEM for the whole
// kernel (M = 64 uses 16 lanes per 32-lane quarter, so two taps interleave in the same columns at lane offsets 0 / 16);
// each persistent CTA reduces its share of the voxels and adds its partial dW to global memory once, at the end.
//
// Stride-2 geometry (GEO = 1; kind 1, and kind 2 = transposed with the roles of x and dz swapped by the caller):
//   dW[tap][ci][co] = sum_{b,od,oh,ow} x[b, 2od+kd-1, 2oh+kh-1, 2ow+kw-1, ci] * dz[b,od,oh,ow,co]
// The x window (17 x 33 fine voxels for an 8 x 16 tile of dz) is de-interleaved into its 4 (row, column) parity sub-planes
// while staging, exactly like the forward stride-2 producer, so that 16 consecutive outputs again read 16 consecutive
// positions of one sub-plane and a tap is a start address.
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"

namespace {

using namespace dpf;

constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kMmaWarps = 4;             // the rows of a tile are dealt round-robin to 4 issuing warps (see conv3d_tc.cu)
constexpr int kProdWarps = 4;
constexpr int kThreads = (kEpiWarps + kMmaWarps + kProdWarps) * 32;
constexpr int kMaxTaps = 27;
constexpr int kWT = 16;                 // tile width (one K=16 segment per row)

struct WgradParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* dz;
  float* dw;                            // [ntaps][CIN][cout] fp32, accumulated with atomics
  int B, D, H, W;                       // x grid
  int Dz, Hz, Wz;                       // dz grid (= x grid for stride 1, ceil(x/2) for stride 2)
  int x_cstride, x_coff, z_cstride, z_coff, cout;
  int ntaps, min_dd;
  int tiles_h, tiles_w, ntiles;
  signed char tap_dd[kMaxTaps];
  short tap_off[kMaxTaps];              // start offset of the tap inside an x slot, in 16-byte positions
};

template <int GEO, int CIN, int NPAD, int NSX, int NSZ>
struct WCfg {
  static constexpr int NCH = CIN / 8;
  static constexpr int TR = GEO == 0 ? 16 : 8;                 // dz tile rows
  static constexpr int XROWS = GEO == 0 ? 18 : 17;             // x window rows / columns loaded
  static constexpr int XCOLS = GEO == 0 ? kWT + 2 : 2 * kWT + 1;
  static constexpr int WP = GEO == 0 ? kWT + 2 : kWT + 1;      // row pitch (positions) of a (sub-)plane
  static constexpr int SUB_POS = (TR + 1) * WP;                // positions of one parity sub-plane (GEO 1)
  static constexpr int XPLANE = (GEO == 0 ? XROWS * WP : 4 * SUB_POS) * 16;
  static constexpr int WANT = (NCH == 4) ? 32 : 16;
  static constexpr int XCH = XPLANE + ((WANT - (XPLANE % 128)) + 128) % 128;     // x chunk-plane pitch (= SBO of A)
  static constexpr int XSLOT = NCH * XCH;
  static constexpr int ZCHUNKS = NPAD / 8;
  static constexpr int ZCH = TR * kWT * 16 + 32;                                  // dz chunk-plane pitch (= SBO of B)
  static constexpr int ZSLOT = ZCHUNKS * ZCH;
  static constexpr int ACC_COLS = ((kMaxTaps + 1) / 2) * NPAD;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  // A reads M = 64 channel rows = 8 chunk planes: with CIN = 32 the upper 4 point past the slot (results discarded),
  // so keep 4 extra chunk planes of addressable shared memory behind the x ring (the dz ring provides them)
  static constexpr int TAIL_WANT = (8 - NCH) * XCH;
  static constexpr int TAIL_PAD = TAIL_WANT > NSZ * ZSLOT ? TAIL_WANT - NSZ * ZSLOT : 0;
  static constexpr int SMEM_BYTES = NSX * XSLOT + NSZ * ZSLOT + TAIL_PAD + (2 * NSX + 2 * NSZ + 1) * 8 + 16 + 128;
  static_assert(ACC_COLS <= 512, "accumulators do not fit TMEM");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

// MN-major, no swizzle: element (mn, k) at start + (mn%8)*2 + (mn/8)*SBO + (k%8)*16 + (k/8)*LBO
__device__ __forceinline__ uint32_t idesc_mn(int m, int n) {
  return umma_idesc_bf16_f32(m, n) | (1u << 15) | (1u << 16);          // a_major = b_major = MN
}

template <int GEO, int CIN, int NPAD, int NSX, int NSZ>
__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  using C = WCfg<GEO, CIN, NPAD, NSX, NSZ>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  uint8_t* s_x = smem;
  uint8_t* s_z = smem + NSX * C::XSLOT;
  uint64_t* bar_xfull = reinterpret_cast<uint64_t*>(s_z + NSZ * C::ZSLOT + C::TAIL_PAD);
  uint64_t* bar_xempty = bar_xfull + NSX;
  uint64_t* bar_zfull = bar_xempty + NSX;
  uint64_t* bar_zempty = bar_zfull + NSZ;
  uint64_t* bar_done = bar_zempty + NSZ;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSX; ++i) { mbar_init(&bar_xfull[i], kProdWarps); mbar_init(&bar_xempty[i], kMmaWarps); }
    for (int i = 0; i < NSZ; ++i) { mbar_init(&bar_zfull[i], kProdWarps); mbar_init(&bar_zempty[i], kMmaWarps); }
    mbar_init(bar_done, kMmaWarps);
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

  if (warp < kEpiWarps) {
    // zero the accumulators (every MMA accumulates), then let the MMA warp start
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c0 = 0; c0 < C::TMEM_COLS; c0 += 16) tmem_zero16(lane_base + c0);
    tmem_st_wait();
    tc_fence_before_sync();
  }
  __syncthreads();
  tc_fence_after_sync();

  if (warp >= kMmaWarp + kMmaWarps) {
    // =================================== producers: x halo windows and dz tiles ===========================
    const int ptid = threadIdx.x - (kMmaWarp + kMmaWarps) * 32;
    constexpr int XPR = C::XCOLS * C::NCH, XPIECES = C::XROWS * XPR;
    constexpr int ZPR = kWT * C::ZCHUNKS, ZPIECES = C::TR * ZPR;
    uint32_t gx_base = 0, gz = 0;
    int prev_x[2] = {-1, -1}, prev_z = -1;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int b = tile / (p.tiles_w * p.tiles_h);
      const int h0 = th * C::TR, w0 = tw * kWT;                          // dz tile origin
      const int xh0 = GEO == 0 ? h0 - 1 : 2 * h0 - 1, xw0 = GEO == 0 ? w0 - 1 : 2 * w0 - 1;
      int xl = 0;                                                         // next x plane of this tile to load
      for (int d = 0; d < p.Dz; ++d, ++gz) {
        int cur_x[2] = {-1, -1}, nx = 0;
        const int x_hi = GEO == 0 ? d : min(2 * d + 1, D - 1);            // highest x plane output plane d needs
        for (; xl <= x_hi; ++xl) {
          const uint32_t gx = gx_base + xl;
          const int sx = gx % NSX;
          mbar_wait(&bar_xempty[sx], ((gx / NSX) & 1u) ^ 1u);
          const uint32_t xb = smem_u32(s_x + sx * C::XSLOT);
          const __nv_bfloat16* xp = p.x + (static_cast<size_t>(b) * D + xl) * H * static_cast<size_t>(W) * p.x_cstride + p.x_coff;
#pragma unroll 4
          for (int q = ptid; q < XPIECES; q += kProdWarps * 32) {
            const int row = q / XPR, rem = q - row * XPR;
            const int col = rem / C::NCH, c8 = rem - col * C::NCH;
            const int h = xh0 + row, w = xw0 + col;
            const bool ok = (h >= 0) && (h < H) && (w >= 0) && (w < W);
            const __nv_bfloat16* src = ok ? (xp + (static_cast<size_t>(h) * W + w) * p.x_cstride + c8 * 8) : p.x;
            const int pos = GEO == 0 ? row * C::WP + col
                                     : ((row & 1) * 2 + (col & 1)) * C::SUB_POS + (row >> 1) * C::WP + (col >> 1);
            cp_async16_zfill(xb + c8 * C::XCH + pos * 16, src, ok);
          }
          cur_x[nx++] = sx;
        }
        const int sz = gz % NSZ;
        mbar_wait(&bar_zempty[sz], ((gz / NSZ) & 1u) ^ 1u);
        const uint32_t zb = smem_u32(s_z + sz * C::ZSLOT);
        const __nv_bfloat16* zp = p.dz + (static_cast<size_t>(b) * p.Dz + d) * p.Hz * static_cast<size_t>(p.Wz) * p.z_cstride + p.z_coff;
#pragma unroll 4
        for (int q = ptid; q < ZPIECES; q += kProdWarps * 32) {
          const int row = q / ZPR, rem = q - row * ZPR;
          const int col = rem / C::ZCHUNKS, c8 = rem - col * C::ZCHUNKS;
          const int h = h0 + row, w = w0 + col;
          const bool ok = (h < p.Hz) && (w < p.Wz) && (c8 * 8 < p.cout);
          const __nv_bfloat16* src = ok ? (zp + (static_cast<size_t>(h) * p.Wz + w) * p.z_cstride + c8 * 8) : p.dz;
          cp_async16_zfill(zb + c8 * C::ZCH + (row * kWT + col) * 16, src, ok);
        }
        cp_async_commit();
        if (GEO == 1) {
          // stride 2 holds 3 x planes per step and loads 2 new ones: with a 4-slot ring the next step's loads cannot be issued
          // before this step has been consumed, so this step is signalled as soon as it has landed (no one-step lag)
          cp_async_wait<0>();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (cur_x[0] >= 0) mbar_arrive(&bar_xfull[cur_x[0]]);
            if (cur_x[1] >= 0) mbar_arrive(&bar_xfull[cur_x[1]]);
            mbar_arrive(&bar_zfull[sz]);
          }
          continue;
        }
        if (prev_z >= 0) {
          cp_async_wait<1>();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (prev_x[0] >= 0) mbar_arrive(&bar_xfull[prev_x[0]]);
            if (prev_x[1] >= 0) mbar_arrive(&bar_xfull[prev_x[1]]);
            mbar_arrive(&bar_zfull[prev_z]);
          }
        }
        prev_x[0] = cur_x[0]; prev_x[1] = cur_x[1]; prev_z = sz;
      }
      gx_base += D;
    }
    if (prev_z >= 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (prev_x[0] >= 0) mbar_arrive(&bar_xfull[prev_x[0]]);
        if (prev_x[1] >= 0) mbar_arrive(&bar_xfull[prev_x[1]]);
        mbar_arrive(&bar_zfull[prev_z]);
      }
    }
  } else if (warp >= kMmaWarp) {
    // =================================== MMA issuers ======================================================
    const int mw = warp - kMmaWarp;
    const uint32_t idesc = idesc_mn(64, NPAD);
    const uint64_t adesc_hi = umma_desc_nosw(0, 128, C::XCH);          // LBO = next 8 positions, SBO = next channel chunk
    const uint64_t bdesc_hi = umma_desc_nosw(0, 128, C::ZCH);
    const uint32_t xbase = smem_u32(s_x) >> 4, zbase = smem_u32(s_z) >> 4;
    const bool leader = elect_one();
    uint32_t gx_base = 0, gz = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int waited = -1;
      for (int d = 0; d < p.Dz; ++d, ++gz) {
        mbar_wait(&bar_zfull[gz % NSZ], (gz / NSZ) & 1u);
        tc_fence_after_sync();
        const uint32_t z0 = zbase + (gz % NSZ) * (C::ZSLOT >> 4);
        for (int t = 0; t < p.ntaps; ++t) {
          const int pin = (GEO == 0 ? d : 2 * d) + p.tap_dd[t];
          if (pin < 0 || pin >= D) continue;
          const uint32_t gx = gx_base + pin;
          if (pin > waited) {
            mbar_wait(&bar_xfull[gx % NSX], (gx / NSX) & 1u);
            tc_fence_after_sync();
            waited = pin;
          }
          const uint32_t a0 = xbase + (gx % NSX) * (C::XSLOT >> 4) + p.tap_off[t];
          const uint32_t acc = tmem_base + (static_cast<uint32_t>((t & 1) * 16) << 16) + (t >> 1) * NPAD;
          if (leader) {
#pragma unroll 4
            for (int row = mw; row < C::TR; row += kMmaWarps) {
              const uint64_t adesc = adesc_hi | static_cast<uint64_t>((a0 + row * C::WP) & 0x3FFF);
              const uint64_t bdesc = bdesc_hi | static_cast<uint64_t>((z0 + row * kWT) & 0x3FFF);
              umma_bf16(acc, adesc, bdesc, idesc, true);
            }
          }
        }
        if (leader) {
          umma_commit(&bar_zempty[gz % NSZ]);
          // release the x planes no later output plane of this tile reads
          const int lo = GEO == 0 ? d + p.min_dd : 2 * d - 1;                     // lowest plane still held for plane d
          const int hi = (d == p.Dz - 1) ? D - 1 : (GEO == 0 ? lo : 2 * d);        // release lo..hi
          for (int q = max(lo, 0); q <= hi; ++q) umma_commit(&bar_xempty[(gx_base + q) % NSX]);
        }
        __syncwarp();
      }
      gx_base += D;
    }
    if (leader) umma_commit(bar_done);
    __syncwarp();
  } else {
    // =================================== epilogue: TMEM partial dW -> global (atomics), once per CTA ======
    mbar_wait(bar_done, 0);
    tc_fence_after_sync();
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int ci = warp * 16 + (lane & 15);                 // M = 64: rows 16q..16q+15 live in lanes 0..15 of quarter q,
    const int todd = lane >> 4;                             // the interleaved (odd) tap in lanes 16..31
    for (int j = 0; j < (p.ntaps + 1) / 2; ++j) {
      const int t = 2 * j + todd;
#pragma unroll
      for (int c0 = 0; c0 < NPAD; c0 += 16) {
        constexpr int CHUNK = NPAD < 16 ? NPAD : 16;
        uint32_t v[16];
        __syncwarp();
        if (NPAD >= 16) tmem_ld16(lane_base + j * NPAD + c0, v);
        else tmem_ld8(lane_base + j * NPAD + c0, v);
        tmem_ld_wait();
        if (t < p.ntaps && ci < CIN) {
#pragma unroll
          for (int k = 0; k < CHUNK; ++k)
            if (c0 + k < p.cout) atomicAdd(p.dw + (static_cast<size_t>(t) * CIN + ci) * p.cout + c0 + k, __uint_as_float(v[k]));
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int GEO, int CIN, int NPAD, int NSX, int NSZ>
int launch_wgrad(WgradParams kp, cudaStream_t st) {
  using C = WCfg<GEO, CIN, NPAD, NSX, NSZ>;
  kp.tiles_h = (kp.Hz + C::TR - 1) / C::TR;
  kp.tiles_w = (kp.Wz + kWT - 1) / kWT;
  kp.ntiles = kp.B * kp.tiles_h * kp.tiles_w;
  for (int t = 0; t < kp.ntaps; ++t) {                         // tap (dh, dw) in 0..2 was left in tap_off as dh * 3 + dw
    const int dh = kp.tap_off[t] / 3, dw = kp.tap_off[t] % 3;
    kp.tap_off[t] = static_cast<short>(GEO == 0 ? dh * C::WP + dw
                                                : ((dh & 1) * 2 + (dw & 1)) * C::SUB_POS + (dh >> 1) * C::WP + (dw >> 1));
  }
  auto kern = wgrad_tc_kernel<GEO, CIN, NPAD, NSX, NSZ>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return dpf::fail("dpf_conv3d_wgrad: cannot opt in to %d B shared memory: %s", C::SMEM_BYTES, cudaGetErrorString(e));
    attr_done = true;
  }
  kern<<<std::min(kp.ntiles, dpf::sm_count()), kThreads, C::SMEM_BYTES, st>>>(kp);
  return dpf::after_launch("dpf_conv3d_wgrad");
}

}  // namespace

extern "C" int dpf_conv3d_wgrad(int kind, const void* x, const void* dz, float* dw, int B, int D, int H, int W, int Cin,
                                int x_cstride, int x_coff, int Cout, int z_cstride, int z_coff, void* stream) {
  DPF_REQUIRE(x && dz && dw, "dpf_conv3d_wgrad: null pointer");
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(dz), "dpf_conv3d_wgrad: pointers must be 16-byte aligned");
  DPF_REQUIRE(kind == 0 || kind == 1 || kind == 3 || kind == 4,
              "dpf_conv3d_wgrad: kind %d not built (0, 3, 4 stride 1; 1 stride 2; the transposed kind 2 is kind 1 with x and dz swapped)", kind);
  DPF_REQUIRE(Cin == 32 || (Cin == 64 && kind != 1), "dpf_conv3d_wgrad: Cin=%d must be 32 or 64 (32 per launch for kind 1)", Cin);
  DPF_REQUIRE(Cout >= 1 && Cout <= 32, "dpf_conv3d_wgrad: Cout=%d must be in [1,32] per launch (split on the host)", Cout);
  DPF_REQUIRE(B > 0 && D > 0 && D <= 64 && H > 0 && W > 0, "dpf_conv3d_wgrad: bad shape");
  DPF_REQUIRE(x_cstride % 8 == 0 && x_coff % 8 == 0 && x_coff + Cin <= x_cstride, "dpf_conv3d_wgrad: bad x channel window");
  DPF_REQUIRE(z_cstride % 8 == 0 && z_coff % 8 == 0 && z_coff + ((Cout + 7) / 8) * 8 <= z_cstride,
              "dpf_conv3d_wgrad: dz channels must be padded to a multiple of 8 (z_cstride=%d, Cout=%d)", z_cstride, Cout);
  WgradParams kp{};
  kp.x = reinterpret_cast<const __nv_bfloat16*>(x);
  kp.dz = reinterpret_cast<const __nv_bfloat16*>(dz);
  kp.dw = dw;
  kp.B = B; kp.D = D; kp.H = H; kp.W = W;
  kp.Dz = kind == 1 ? (D + 1) / 2 : D; kp.Hz = kind == 1 ? (H + 1) / 2 : H; kp.Wz = kind == 1 ? (W + 1) / 2 : W;
  kp.x_cstride = x_cstride; kp.x_coff = x_coff; kp.z_cstride = z_cstride; kp.z_coff = z_coff; kp.cout = Cout;
  int t = 0;
  if (kind == 0 || kind == 1) {
    for (int kd = 0; kd < 3; ++kd)
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw, ++t) { kp.tap_dd[t] = kd - 1; kp.tap_off[t] = static_cast<short>(kh * 3 + kw); }
    kp.min_dd = -1;
  } else if (kind == 3) {
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw, ++t) { kp.tap_dd[t] = 0; kp.tap_off[t] = static_cast<short>(kh * 3 + kw); }
    kp.min_dd = 0;
  } else {
    kp.tap_dd[0] = 0; kp.tap_off[0] = 1 * 3 + 1; t = 1; kp.min_dd = 0;
  }
  kp.ntaps = t;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int npad = Cout <= 8 ? 8 : (Cout <= 16 ? 16 : 32);
  if (kind == 1) return launch_wgrad<1, 32, 32, 4, 3>(kp, st);
  if (Cin == 32 && npad == 32) return launch_wgrad<0, 32, 32, 5, 3>(kp, st);
  if (Cin == 32 && npad == 16) return launch_wgrad<0, 32, 16, 5, 3>(kp, st);
  if (Cin == 32 && npad == 8) return launch_wgrad<0, 32, 8, 5, 3>(kp, st);
  if (Cin == 64 && npad == 32) return launch_wgrad<0, 64, 32, 4, 3>(kp, st);
  if (Cin == 64 && npad == 16) return launch_wgrad<0, 64, 16, 4, 3>(kp, st);
  if (Cin == 64 && npad == 8) return launch_wgrad<0, 64, 8, 4, 3>(kp, st);
  return dpf::fail("dpf_conv3d_wgrad: no kernel for Cin=%d Cout=%d", Cin, Cout);
}
