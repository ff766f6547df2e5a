// 3-D convolution as an implicit GEMM on tcgen05 tensor cores (sm_100a), BN / bias / residual / ReLU fused.
//
// Replaces nn.Conv3d + BatchNorm3d (+ReLU)(+add) of the reference's PSMNetHGAggregation
// (src/model/stereodpnet/modules.py:204-337; convbn_3d in src/module/asm/basics.py:32-36) and the mask convolutions of
// MaskingAttention (src/module/asm/asm.py:141-146).
//
// GEMM view per tap: D[128 voxels x Cout] += A[128 voxels x Cin] * W_tap[Cin x Cout], 27 taps (x Cin/16 k-steps)
// accumulate into one TMEM tile.  Activations are bf16 NDHWC.
//
// Shared-memory layout (the point of the design).  An output block is 16 rows (h) x 8 columns (w) of ONE depth
// plane = 128 GEMM rows.  Input planes are staged per (tile, depth plane) as a halo window of 18 x (WT+2) voxels,
// stored channel-chunk-planar:   slot[chunk c8][row][col][8 channels] (16 B per voxel per chunk).  In the UMMA
// no-swizzle K-major canonical layout a row of the A operand is 16 B, 8 consecutive rows are contiguous (= 8
// consecutive w), and consecutive 8-row groups are SBO apart (= the window row pitch), k-chunks are LBO apart
// (= the chunk-plane pitch).  Every one of the 27 taps is then just a different START ADDRESS into the same
// staged window: no im2col, no per-tap copies, no padding waste -- each input voxel is loaded once per tile
// (plus halo) with 16-byte LDGSTS (zero-filled outside the image, which implements the conv padding), and is
// read 27 times by the tensor core straight from shared memory.
//
// Pipeline: 4 producer warps (cp.async) -> ring of NS plane slots (full/empty mbarriers) -> 1 MMA thread
// (tcgen05.mma, accumulators double-buffered in TMEM) -> 4 epilogue warps (tcgen05.ld, scale/shift, residual, ReLU,
// bf16 pack, 128-bit stores).  Persistent CTAs, one per SM; weights of all taps stay resident in shared memory.
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"

namespace {

using namespace dpf;

constexpr int kEpiWarps = 4;
constexpr int kProdWarps = 4;
constexpr int kMmaWarp = kEpiWarps;                                   // warp 4
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;          // 288
constexpr int kRows = 18;                                             // 16 output rows + halo
constexpr int kMaxTaps = 27;

struct ConvKParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
  void* y;
  const float* scale;
  const float* shift;
  const void* residual;
  float* stats;
  int B, D, H, W;
  int cout, y_f32, y_cstride, y_coff, relu;
  int ntaps, min_dd;
  int tiles_h, tiles_w, ntiles;
  signed char tap_dd[kMaxTaps];
  signed char tap_dh[kMaxTaps];
  signed char tap_dw[kMaxTaps];
};

template <int CIN, int NPAD, int WT, int NS>
struct Cfg {
  static constexpr int NCH = CIN / 8;
  static constexpr int WP = WT + 2;
  static constexpr int PLANE_BYTES = kRows * WP * 16;
  // chunk-plane pitch: 16-B multiple whose residue mod 128 spreads a quarter-warp's cp.async writes over all banks
  static constexpr int WANT = (NCH == 4) ? 32 : 16;
  static constexpr int CH_STRIDE = PLANE_BYTES + ((WANT - (PLANE_BYTES % 128)) + 128) % 128;
  static constexpr int SLOT_BYTES = NCH * CH_STRIDE;
  static constexpr int W_TAP_BYTES = NCH * NPAD * 16;
  static constexpr int W_BYTES = kMaxTaps * W_TAP_BYTES;
  static constexpr int NBLK = WT / 8;
  static constexpr int ACC_COLS = NBLK * NPAD;
  static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
  static constexpr int KSTEPS = CIN / 16;
  static constexpr int SMEM_BYTES = W_BYTES + NS * SLOT_BYTES + 2 * NPAD * 4 + (2 * NS + 4) * 8 + 16 + 128;
  static_assert(2 * ACC_COLS <= 512, "accumulators do not fit TMEM");
  static_assert(CH_STRIDE % 16 == 0, "chunk pitch must be a 16-byte multiple");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

template <int CIN, int NPAD, int WT, int NS>
__global__ void __launch_bounds__(kThreads, 1) conv3d_tc_kernel(const __grid_constant__ ConvKParams p) {
  using C = Cfg<CIN, NPAD, WT, NS>;
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

  // ---- one-time setup ---------------------------------------------------------------------------------
  {
    const int nbytes = p.ntaps * C::W_TAP_BYTES;
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = threadIdx.x; i < nbytes / 16; i += kThreads) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < NPAD; i += kThreads) {
      s_scale[i] = (p.scale != nullptr && i < p.cout) ? p.scale[i] : 1.0f;
      s_shift[i] = (p.shift != nullptr && i < p.cout) ? p.shift[i] : 0.0f;
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

  const int D = p.D, H = p.H, W = p.W;

  if (warp > kMmaWarp) {
    // =================================== producers: global -> shared ring ===================================
    const int ptid = threadIdx.x - (kMmaWarp + 1) * 32;          // 0..127
    constexpr int PIECES_PER_ROW = C::WP * C::NCH;
    constexpr int PIECES = kRows * PIECES_PER_ROW;
    uint32_t g = 0;
    int prev_slot = -1;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int b = tile / (p.tiles_w * p.tiles_h);
      const int h0 = th * 16 - 1, w0 = tw * WT - 1;              // window origin (with halo)
      for (int pl = 0; pl < D; ++pl, ++g) {
        const int slot = g % NS;
        const uint32_t ph = (g / NS) & 1u;
        mbar_wait(&bar_empty[slot], ph ^ 1u);
        const uint32_t sbase = smem_u32(s_slots + slot * C::SLOT_BYTES);
        const __nv_bfloat16* xplane = p.x + (static_cast<size_t>(b) * D + pl) * H * static_cast<size_t>(W) * CIN;
#pragma unroll 4
        for (int q = ptid; q < PIECES; q += kProdWarps * 32) {
          const int row = q / PIECES_PER_ROW;
          const int rem = q - row * PIECES_PER_ROW;
          const int col = rem / C::NCH;
          const int c8 = rem - col * C::NCH;
          const int h = h0 + row, w = w0 + col;
          const bool ok = (h >= 0) && (h < H) && (w >= 0) && (w < W);
          const __nv_bfloat16* src = ok ? (xplane + (static_cast<size_t>(h) * W + w) * CIN + c8 * 8) : p.x;
          cp_async16_zfill(sbase + c8 * C::CH_STRIDE + (row * C::WP + col) * 16, src, ok);
        }
        cp_async_commit();
        if (prev_slot >= 0) {                                    // complete the previous plane (one group of lag)
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
    // =================================== MMA issuer (one thread) ==========================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(128, NPAD);
      const uint32_t wbase = smem_u32(s_w);
      uint32_t g_base = 0, it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int tw = tile % p.tiles_w;
        const int nblk = min(C::NBLK, (W - tw * WT + 7) >> 3);
        int waited = -1;
        for (int d = 0; d < D; ++d, ++it) {
          const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
          mbar_wait(&bar_tempty[as], aph ^ 1u);
          tc_fence_after_sync();
          bool first = true;
          for (int t = 0; t < p.ntaps; ++t) {
            const int pin = d + p.tap_dd[t];
            if (pin < 0 || pin >= D) continue;
            const uint32_t gp = g_base + pin;
            const uint32_t slot = gp % NS;
            if (pin > waited) {
              mbar_wait(&bar_full[slot], (gp / NS) & 1u);
              tc_fence_after_sync();
              waited = pin;
            }
            const uint32_t a0 = smem_u32(s_slots + slot * C::SLOT_BYTES) + (p.tap_dh[t] * C::WP + p.tap_dw[t]) * 16;
            const uint32_t b0 = wbase + t * C::W_TAP_BYTES;
#pragma unroll
            for (int ks = 0; ks < C::KSTEPS; ++ks) {
              const uint64_t bdesc = umma_desc_nosw(b0 + ks * 2 * NPAD * 16, NPAD * 16, 128);
              for (int blk = 0; blk < nblk; ++blk) {
                const uint64_t adesc = umma_desc_nosw(a0 + ks * 2 * C::CH_STRIDE + blk * 128, C::CH_STRIDE, C::WP * 16);
                umma_bf16(tmem_base + (as * C::NBLK + blk) * NPAD, adesc, bdesc, idesc, !(first && ks == 0));
              }
            }
            first = false;
          }
          umma_commit(&bar_tfull[as]);
          // input planes no later output plane of this tile needs
          const int rel = d + p.min_dd;
          if (d == D - 1) {
            for (int q = max(rel, 0); q < D; ++q) umma_commit(&bar_empty[(g_base + q) % NS]);
          } else if (rel >= 0) {
            umma_commit(&bar_empty[(g_base + rel) % NS]);
          }
        }
        g_base += D;
      }
    }
    __syncwarp();
  } else {
    // =================================== epilogue: TMEM -> registers -> global ============================
    uint32_t it = 0;
    const int m = warp * 32 + lane;
    const int hrow = m >> 3, wcol = m & 7;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int b = tile / (p.tiles_w * p.tiles_h);
      const int nblk = min(C::NBLK, (W - tw * WT + 7) >> 3);
      const int h = th * 16 + hrow;
      for (int d = 0; d < D; ++d, ++it) {
        const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
        mbar_wait(&bar_tfull[as], aph);
        tc_fence_after_sync();
        for (int blk = 0; blk < nblk; ++blk) {
          const int w = tw * WT + blk * 8 + wcol;
          const bool ok = (h < H) && (w < W);
          const size_t vox = ((static_cast<size_t>(b) * D + d) * H + h) * static_cast<size_t>(W) + w;
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + (as * C::NBLK + blk) * NPAD;
#pragma unroll
          for (int c0 = 0; c0 < NPAD; c0 += 16) {
            uint32_t v[16];
            __syncwarp();                                        // tcgen05.ld is .sync.aligned: keep the warp converged
            tmem_ld16(taddr + c0, v);
            tmem_ld_wait();
            if (!ok || c0 >= p.cout) continue;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * s_scale[c0 + j] + s_shift[c0 + j];
            if (p.y_f32) {
              float* yo = reinterpret_cast<float*>(p.y) + vox * p.y_cstride + p.y_coff + c0;
              const float* ro = p.residual ? reinterpret_cast<const float*>(p.residual) + vox * p.y_cstride + p.y_coff + c0 : nullptr;
              const int n = min(16, p.cout - c0);
              for (int j = 0; j < n; ++j) {
                float val = f[j] + (ro ? ro[j] : 0.f);
                if (p.relu) val = fmaxf(val, 0.f);
                yo[j] = val;
              }
            } else {
              __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(p.y) + vox * p.y_cstride + p.y_coff + c0;
              const int n = min(16, p.cout - c0);                // multiple of 8 (checked on the host)
              if (p.residual) {
                const __nv_bfloat16* ro = reinterpret_cast<const __nv_bfloat16*>(p.residual) + vox * p.y_cstride + p.y_coff + c0;
                for (int j8 = 0; j8 < n; j8 += 8) {
                  const uint4 r = *reinterpret_cast<const uint4*>(ro + j8);
                  f[j8 + 0] += bf16_lo(r.x); f[j8 + 1] += bf16_hi(r.x);
                  f[j8 + 2] += bf16_lo(r.y); f[j8 + 3] += bf16_hi(r.y);
                  f[j8 + 4] += bf16_lo(r.z); f[j8 + 5] += bf16_hi(r.z);
                  f[j8 + 6] += bf16_lo(r.w); f[j8 + 7] += bf16_hi(r.w);
                }
              }
              if (p.relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
              }
              for (int j8 = 0; j8 < n; j8 += 8) {
                uint4 o;
                o.x = pack_bf16x2(f[j8 + 0], f[j8 + 1]);
                o.y = pack_bf16x2(f[j8 + 2], f[j8 + 3]);
                o.z = pack_bf16x2(f[j8 + 4], f[j8 + 5]);
                o.w = pack_bf16x2(f[j8 + 6], f[j8 + 7]);
                *reinterpret_cast<uint4*>(yo + j8) = o;
              }
            }
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[as]);
      }
    }
  }

  // ---- teardown ------------------------------------------------------------------------------------------
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int CIN, int NPAD, int WT, int NS>
int launch(const ConvKParams& kp_in, cudaStream_t st) {
  using C = Cfg<CIN, NPAD, WT, NS>;
  ConvKParams kp = kp_in;
  kp.tiles_h = (kp.H + 15) / 16;
  kp.tiles_w = (kp.W + WT - 1) / WT;
  kp.ntiles = kp.B * kp.tiles_h * kp.tiles_w;
  auto kern = conv3d_tc_kernel<CIN, NPAD, WT, NS>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return dpf::fail("dpf_conv3d_fwd: cannot opt in to %d B shared memory: %s", C::SMEM_BYTES, cudaGetErrorString(e));
    attr_done = true;
  }
  const int grid = std::min(kp.ntiles, dpf::sm_count());
  kern<<<grid, kThreads, C::SMEM_BYTES, st>>>(kp);
  return dpf::after_launch("dpf_conv3d_fwd");
}

int npad_for(int cout) { return cout <= 16 ? 16 : (cout <= 32 ? 32 : 64); }

}  // namespace

extern "C" long long dpf_conv3d_weight_elems(int kind, int Cin, int Cout) {
  const int ntaps = (kind == 0 || kind == 1 || kind == 2) ? 27 : (kind == 3 ? 9 : 1);
  return static_cast<long long>(ntaps) * Cin * npad_for(Cout);
}

extern "C" int dpf_conv3d_fwd(const dpf_conv3d_args* a, void* stream) {
  DPF_REQUIRE(a != nullptr, "dpf_conv3d_fwd: null args");
  DPF_REQUIRE(a->x && a->w && a->y, "dpf_conv3d_fwd: null tensor pointer");
  DPF_REQUIRE(DPF_ALIGNED16(a->x) && DPF_ALIGNED16(a->w) && DPF_ALIGNED16(a->y), "dpf_conv3d_fwd: pointers must be 16-byte aligned");
  DPF_REQUIRE(a->kind == 0 || a->kind == 3 || a->kind == 4, "dpf_conv3d_fwd: kind %d not built yet (0, 3, 4 are)", a->kind);
  DPF_REQUIRE(a->Cin == 32 || a->Cin == 64, "dpf_conv3d_fwd: Cin=%d must be 32 or 64", a->Cin);
  DPF_REQUIRE(a->Cout >= 1 && a->Cout <= 64, "dpf_conv3d_fwd: Cout=%d must be in [1,64] (split wider layers on the host)", a->Cout);
  DPF_REQUIRE(a->Cin == 32 || a->Cout <= 32, "dpf_conv3d_fwd: Cin=64 supports Cout<=32 per launch (split on the host)");
  DPF_REQUIRE(a->B > 0 && a->D > 0 && a->D <= 64 && a->H > 0 && a->W > 0, "dpf_conv3d_fwd: bad shape");
  DPF_REQUIRE(a->y_f32 || (a->Cout % 8 == 0 && a->y_cstride % 8 == 0 && a->y_coff % 8 == 0),
              "dpf_conv3d_fwd: bf16 output needs Cout, y_cstride, y_coff multiples of 8");
  DPF_REQUIRE(a->y_cstride >= a->y_coff + a->Cout, "dpf_conv3d_fwd: y_cstride too small");
  ConvKParams kp{};
  kp.x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  kp.w = reinterpret_cast<const __nv_bfloat16*>(a->w);
  kp.y = a->y;
  kp.scale = a->scale;
  kp.shift = a->shift;
  kp.residual = a->residual;
  kp.stats = a->stats;
  kp.B = a->B; kp.D = a->D; kp.H = a->H; kp.W = a->W;
  kp.cout = a->Cout; kp.y_f32 = a->y_f32; kp.y_cstride = a->y_cstride; kp.y_coff = a->y_coff; kp.relu = a->relu;
  int t = 0;
  if (a->kind == 0) {
    for (int kd = 0; kd < 3; ++kd)
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw, ++t) {
          kp.tap_dd[t] = static_cast<signed char>(kd - 1);
          kp.tap_dh[t] = static_cast<signed char>(kh);
          kp.tap_dw[t] = static_cast<signed char>(kw);
        }
    kp.min_dd = -1;
  } else if (a->kind == 3) {
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw, ++t) {
        kp.tap_dd[t] = 0;
        kp.tap_dh[t] = static_cast<signed char>(kh);
        kp.tap_dw[t] = static_cast<signed char>(kw);
      }
    kp.min_dd = 0;
  } else {
    kp.tap_dd[0] = 0; kp.tap_dh[0] = 1; kp.tap_dw[0] = 1;
    t = 1;
    kp.min_dd = 0;
  }
  kp.ntaps = t;
  DPF_REQUIRE(a->stats == nullptr, "dpf_conv3d_fwd: fused batch statistics are not built yet");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int npad = npad_for(a->Cout);
  if (a->Cin == 32 && npad == 32) return launch<32, 32, 24, 5>(kp, st);
  if (a->Cin == 32 && npad == 16) return launch<32, 16, 24, 5>(kp, st);
  if (a->Cin == 32 && npad == 64) return launch<32, 64, 8, 6>(kp, st);
  if (a->Cin == 64 && npad == 32) return launch<64, 32, 8, 5>(kp, st);
  if (a->Cin == 64 && npad == 16) return launch<64, 16, 8, 5>(kp, st);
  return dpf::fail("dpf_conv3d_fwd: no kernel for Cin=%d Cout=%d", a->Cin, a->Cout);
}
