// Stride-2 3x3x3 convolution (pad 1) + BN + ReLU as a PLANE-STREAMED implicit GEMM on tcgen05 (sm_100a): Cin in {32, 64},
// up to 64 output channels in ONE launch.
//
// Replaces nn.Conv3d(k3, s2, p1) + BatchNorm3d + ReLU of the hourglass (dresK.conv1 32->64 and dresK.conv3 64->64,
// src/model/stereodpnet/modules.py:208,213; convbn_3d in src/module/asm/basics.py:32-36).  Round 1 ran them on the generic
// kernel of conv3d_tc.cu with at most 32 input and 32 output channels per launch: 2 launches for conv1, 4 for conv3 (chained
// through an fp32 partial tensor), each re-reading its input, at 14 % of the tensor peak (profiles/r01_conv_s2_32to64.txt).
//
// Here every input plane is staged ONCE per tile and consumed in a single pass (as in the kd-fused stride-1 kernel): input plane p
// feeds output plane p/2 through the kd = 1 taps when p is even, and output planes (p-1)/2 and (p+1)/2 through the kd = 2 and
// kd = 0 taps when p is odd.  The accumulators of consecutive output planes sit in a ring of R TMEM stages; N = Npad = all output
// channels (64 for conv1), so the A window is read once per tap for all of them.  Input channels are consumed in 32-channel
// windows ("k-parts", one shared-memory slot each) that accumulate into the same TMEM tile, which lets the resident weights of all
// 27 taps (110 KB) and a 3-slot ring fit shared memory for Cin = 64 too.
// Window layout: as GEO_S2 of conv3d_tc.cu -- the 33 x 17 input window of a 16 x 8 output block is de-interleaved into its four
// (row, column) parity sub-planes, so that the rows 2*ho + kh - 1 of consecutive outputs are consecutive again and a tap is a
// descriptor start address: sub-plane (kh & 1, kw & 1), offset (kh >> 1, kw >> 1).
// One MMA-issuing thread, fixed order: deterministic.  4 producer warps (cp.async), 1 MMA warp, 4 epilogue warps; persistent CTAs.
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"
#include <algorithm>

namespace {

using namespace dpf;

constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kProdWarps = 4;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;   // 288
constexpr int kWT = 8;                                         // output columns per tile (one 128-row GEMM block)

struct S2Params {
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
  int B, D, H, W, Do, Ho, Wo;
  int x_cstride, x_coff, y_cstride, y_coff, cout, relu;
  int tiles_h, tiles_w, ntiles;
};

template <int CIN, int NPAD, int NS, int R>
struct S2Cfg {
  static constexpr int KPART = 32, NCHP = 4, KPARTS = CIN / KPART, KSTEPS = 2;
  static constexpr int RWIN = 33, CWIN = 2 * kWT + 1;           // input window of a 16 x 8 output block
  static constexpr int RS = 17, WPS = kWT + 1;                  // rows / row pitch of one parity sub-plane
  static constexpr int SUB_POS = RS * WPS;
  static constexpr int PLANE_BYTES = 4 * SUB_POS * 16;
  static constexpr int CH_STRIDE = PLANE_BYTES + ((32 - (PLANE_BYTES % 128)) + 128) % 128;
  static constexpr int SLOT_BYTES = NCHP * CH_STRIDE;
  static constexpr int W_TAP_BYTES = (CIN / 8) * NPAD * 16;     // [c8][NPAD][8]
  static constexpr int W_BYTES = 27 * W_TAP_BYTES;
  static constexpr int COLS = R * NPAD;
  static constexpr int TMEM_COLS = COLS <= 32 ? 32 : COLS <= 64 ? 64 : COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = W_BYTES + NS * SLOT_BYTES + 2 * NPAD * 4 + (2 * NS + 2 * R) * 8 + 16 + 128;
  static_assert(CIN % KPART == 0, "input channels are consumed in 32-channel windows");
  static_assert(COLS <= 512 && R >= 3, "a ring of >= 3 output-plane accumulators must fit TMEM");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

template <int CIN, int NPAD, int NS, int R>
__global__ void __launch_bounds__(kThreads, 1) conv3d_s2_kernel(const __grid_constant__ S2Params p) {
  using C = S2Cfg<CIN, NPAD, NS, R>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  uint8_t* s_w = smem;
  uint8_t* s_slots = smem + C::W_BYTES;
  float* s_scale = reinterpret_cast<float*>(s_slots + NS * C::SLOT_BYTES);
  float* s_shift = s_scale + NPAD;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_shift + NPAD);
  uint64_t* bar_empty = bar_full + NS;
  uint64_t* bar_tfull = bar_empty + NS;
  uint64_t* bar_tempty = bar_tfull + R;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_tempty + R);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = threadIdx.x; i < C::W_BYTES / 16; i += kThreads) dst[i] = __ldg(src + i);
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
    for (int i = 0; i < R; ++i) {
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
  const int D = p.D, H = p.H, W = p.W, Do = p.Do;

  if (warp > kMmaWarp) {
    // =================================== producers: (tile, input plane, k-part) -> slot ring ======================
    const int pwarp = warp - (kMmaWarp + 1);
    constexpr int PIECES_PER_ROW = C::CWIN * C::NCHP;
    uint32_t g = 0;
    int prev_slot = -1;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int b = tile / (p.tiles_w * p.tiles_h);
      const int h0 = th * 32 - 1, w0 = tw * 2 * kWT - 1;
      for (int pl = 0; pl < D; ++pl) {
        const __nv_bfloat16* xplane = p.x + (static_cast<size_t>(b) * D + pl) * H * static_cast<size_t>(W) * p.x_cstride + p.x_coff;
        for (int kp = 0; kp < C::KPARTS; ++kp, ++g) {
          const int slot = g % NS;
          mbar_wait(&bar_empty[slot], ((g / NS) & 1u) ^ 1u);
          const uint32_t sbase = smem_u32(s_slots + slot * C::SLOT_BYTES);
          for (int row = pwarp; row < C::RWIN; row += kProdWarps) {
            const int h = h0 + row;
            const bool hok = (h >= 0) && (h < H);
            const __nv_bfloat16* xrow = xplane + static_cast<size_t>(hok ? h : 0) * W * p.x_cstride + kp * C::KPART;
            const int rowpos = (row & 1) * 2 * C::SUB_POS + (row >> 1) * C::WPS;       // parity sub-plane of this row
#pragma unroll
            for (int q = lane; q < PIECES_PER_ROW; q += 32) {
              const int col = q >> 2, c8 = q & 3;
              const int w = w0 + col;
              const bool ok = hok && (w >= 0) && (w < W);
              const __nv_bfloat16* src = ok ? (xrow + static_cast<size_t>(w) * p.x_cstride + c8 * 8) : p.x;
              const int pos = rowpos + (col & 1) * C::SUB_POS + (col >> 1);
              cp_async16_zfill(sbase + c8 * C::CH_STRIDE + pos * 16, src, ok);
            }
          }
          cp_async_commit();
          if (prev_slot >= 0) {
            cp_async_wait<1>();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[prev_slot]);
          }
          prev_slot = slot;
        }
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
    const uint64_t adesc_hi = umma_desc_nosw(0, C::CH_STRIDE, C::WPS * 16);
    const uint64_t bdesc_hi = umma_desc_nosw(0, NPAD * 16, 128);
    const bool leader = elect_one();
    uint32_t g = 0, q_base = 0;                                    // slot counter; output planes completed before this tile
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      for (int pl = 0; pl < D; ++pl) {
        // output planes fed by input plane pl: even -> (kd 1, pl/2); odd -> (kd 2, (pl-1)/2) and (kd 0, (pl+1)/2)
        const int nsets = (pl & 1) ? (((pl + 1) >> 1) < Do ? 2 : 1) : 1;
        const int kd_a = (pl & 1) ? 2 : 1, do_a = pl >> 1;
        const int do_b = (pl + 1) >> 1;                            // second set (odd planes): kd 0
        // a stage is touched for the first time by its kd = 0 taps (or by plane 0): it must have been drained
        if (pl == 0) {
          const uint32_t Q = q_base;
          mbar_wait(&bar_tempty[Q % R], ((Q / R) & 1u) ^ 1u);
        }
        if (nsets == 2) {
          const uint32_t Q = q_base + do_b;
          mbar_wait(&bar_tempty[Q % R], ((Q / R) & 1u) ^ 1u);
        }
        tc_fence_after_sync();
        for (int kp = 0; kp < C::KPARTS; ++kp, ++g) {
          const uint32_t slot = g % NS;
          mbar_wait(&bar_full[slot], (g / NS) & 1u);
          tc_fence_after_sync();
          if (leader) {
            const uint32_t a_slot = (sbase0 + slot * C::SLOT_BYTES) >> 4;
#pragma unroll 1
            for (int set = 0; set < nsets; ++set) {
              const int kd = set == 0 ? kd_a : 0;
              const uint32_t Q = q_base + (set == 0 ? do_a : do_b);
              const uint32_t acc = tmem_base + (Q % R) * NPAD;
              const bool fresh = (kd == 0) || (pl == 0);           // first contribution to this output plane: overwrite
#pragma unroll
              for (int t9 = 0; t9 < 9; ++t9) {
                const int kh = t9 / 3, kw = t9 % 3;
                const uint32_t a0 = a_slot + ((kh & 1) * 2 + (kw & 1)) * C::SUB_POS + (kh >> 1) * C::WPS + (kw >> 1);
                const uint32_t b0 = wbase + (kd * 9 + t9) * (C::W_TAP_BYTES >> 4) + kp * C::NCHP * NPAD;
#pragma unroll
                for (int ks = 0; ks < C::KSTEPS; ++ks) {
                  const uint64_t adesc = adesc_hi | static_cast<uint64_t>(a0 + ks * 2 * (C::CH_STRIDE >> 4));
                  const uint64_t bdesc = bdesc_hi | static_cast<uint64_t>(b0 + ks * 2 * NPAD);
                  umma_bf16(acc, adesc, bdesc, idesc, !(fresh && kp == 0 && t9 == 0 && ks == 0));
                }
              }
            }
            umma_commit(&bar_empty[slot]);
          }
          __syncwarp();
        }
        // output plane do_a is complete after its kd = 2 plane (odd pl), or after the last plane when D is odd
        if (leader && ((pl & 1) || pl == D - 1)) umma_commit(&bar_tfull[(q_base + do_a) % R]);
        __syncwarp();
      }
      q_base += Do;
    }
  } else {
    // =================================== epilogue: TMEM -> registers -> global ====================================
    uint32_t q_base = 0;
    const int m = warp * 32 + lane;
    const int hrow = m >> 3, wcol = m & 7;
    const bool wide = ((p.y_cstride | p.y_coff) & 15) == 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int b = tile / (p.tiles_w * p.tiles_h);
      const int oh = th * 16 + hrow, ow = tw * kWT + wcol;
      const bool ok = (oh < p.Ho) && (ow < p.Wo);
      for (int od = 0; od < Do; ++od) {
        const uint32_t Q = q_base + od;
        const uint32_t st = Q % R;
        mbar_wait(&bar_tfull[st], (Q / R) & 1u);
        tc_fence_after_sync();
        const size_t vox = ((static_cast<size_t>(b) * Do + od) * p.Ho + (ok ? oh : 0)) * static_cast<size_t>(p.Wo) + (ok ? ow : 0);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + st * NPAD;
#pragma unroll
        for (int c0 = 0; c0 < NPAD; c0 += 16) {
          uint32_t v[16];
          __syncwarp();
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (!ok || c0 >= p.cout) continue;
          const int nst = min(16, p.cout - c0);
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            f[j] = __uint_as_float(v[j]) * s_scale[c0 + j] + s_shift[c0 + j];
            if (p.relu) f[j] = fmaxf(f[j], 0.f);
          }
          uint4 o0, o1;
          o0.x = pack_bf16x2(f[0], f[1]); o0.y = pack_bf16x2(f[2], f[3]); o0.z = pack_bf16x2(f[4], f[5]); o0.w = pack_bf16x2(f[6], f[7]);
          o1.x = pack_bf16x2(f[8], f[9]); o1.y = pack_bf16x2(f[10], f[11]); o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
          __nv_bfloat16* yo = p.y + vox * p.y_cstride + p.y_coff + c0;
          if (wide && nst == 16) st_global_v8(yo, o0, o1);
          else {
            *reinterpret_cast<uint4*>(yo) = o0;
            if (nst > 8) *reinterpret_cast<uint4*>(yo + 8) = o1;
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[st]);
      }
      q_base += Do;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int CIN, int NPAD, int NS, int R>
int launch_s2(S2Params kp, cudaStream_t st) {
  using C = S2Cfg<CIN, NPAD, NS, R>;
  auto kern = conv3d_s2_kernel<CIN, NPAD, NS, R>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return dpf::fail("dpf_conv3d_s2_fwd: cannot opt in to %d B shared memory: %s", C::SMEM_BYTES, cudaGetErrorString(e));
    attr_done = true;
  }
  const int grid = std::min(kp.ntiles, dpf::sm_count());
  kern<<<grid, kThreads, C::SMEM_BYTES, st>>>(kp);
  return dpf::after_launch("dpf_conv3d_s2_fwd");
}

}  // namespace

extern "C" int dpf_conv3d_s2_fwd(const void* x, const void* w, void* y, const float* scale, const float* shift, int B, int D, int H,
                                 int W, int Cin, int Cout, int x_cstride, int x_coff, int y_cstride, int y_coff, int relu, void* stream) {
  DPF_REQUIRE(x && w && y, "dpf_conv3d_s2_fwd: null tensor pointer");
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(w) && DPF_ALIGNED16(y), "dpf_conv3d_s2_fwd: pointers must be 16-byte aligned");
  DPF_REQUIRE(B > 0 && D > 0 && D <= 64 && H > 0 && W > 0, "dpf_conv3d_s2_fwd: bad shape");
  DPF_REQUIRE(Cout >= 8 && Cout % 8 == 0 && Cout <= 64, "dpf_conv3d_s2_fwd: Cout=%d must be a multiple of 8 in [8, 64]", Cout);
  DPF_REQUIRE(x_cstride % 8 == 0 && x_coff % 8 == 0 && x_coff + Cin <= x_cstride, "dpf_conv3d_s2_fwd: bad input channel window");
  DPF_REQUIRE(y_cstride % 8 == 0 && y_coff % 8 == 0 && y_coff + Cout <= y_cstride, "dpf_conv3d_s2_fwd: bad output channel window");
  S2Params kp{};
  kp.x = reinterpret_cast<const __nv_bfloat16*>(x);
  kp.w = reinterpret_cast<const __nv_bfloat16*>(w);
  kp.y = reinterpret_cast<__nv_bfloat16*>(y);
  kp.scale = scale; kp.shift = shift;
  kp.B = B; kp.D = D; kp.H = H; kp.W = W;
  kp.Do = (D + 1) / 2; kp.Ho = (H + 1) / 2; kp.Wo = (W + 1) / 2;
  kp.x_cstride = x_cstride; kp.x_coff = x_coff; kp.y_cstride = y_cstride; kp.y_coff = y_coff; kp.cout = Cout; kp.relu = relu;
  kp.tiles_h = (kp.Ho + 15) / 16;
  kp.tiles_w = (kp.Wo + kWT - 1) / kWT;
  kp.ntiles = B * kp.tiles_h * kp.tiles_w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int npad = Cout <= 16 ? 16 : (Cout <= 32 ? 32 : 64);           // = the packing of dpf_conv3d_fwd (pack_conv_weight)
  if (Cin == 32 && npad == 64) return launch_s2<32, 64, 3, 4>(kp, st);
  if (Cin == 32 && npad == 32) return launch_s2<32, 32, 4, 4>(kp, st);
  if (Cin == 32 && npad == 16) return launch_s2<32, 16, 4, 4>(kp, st);
  if (Cin == 64 && npad == 16) return launch_s2<64, 16, 4, 4>(kp, st);
  if (Cin == 64 && npad == 32) return launch_s2<64, 32, 3, 4>(kp, st);
  return dpf::fail("dpf_conv3d_s2_fwd: no kernel for Cin=%d Cout=%d (built: Cin 32 -> Cout <= 64, Cin 64 -> Cout <= 32 per launch)", Cin, Cout);
}
