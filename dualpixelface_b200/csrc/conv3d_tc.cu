// 3-D convolution as an implicit GEMM on tcgen05 tensor cores (sm_100a), BN / bias / residual / ReLU fused.
//
// Replaces nn.Conv3d / nn.ConvTranspose3d + BatchNorm3d (+ReLU)(+add) of the reference's PSMNetHGAggregation
// (src/model/stereodpnet/modules.py:204-337; convbn_3d in src/module/asm/basics.py:32-36) and the mask convolutions of
// MaskingAttention (src/module/asm/asm.py:141-146).
//
// GEMM view per tap: D[128 voxels x Cout] += A[128 voxels x Cin] * W_tap[Cin x Cout]; taps (x Cin/16 k-steps)
// accumulate into one TMEM tile.  Activations are bf16 NDHWC.
//
// Shared-memory layout (the point of the design).  A GEMM block is 16 rows (h) x 8 columns (w) of ONE depth plane
// = 128 GEMM rows.  Input planes are staged per (tile, depth plane) as a halo window, stored channel-chunk-planar:
//   slot[chunk c8][sub-plane][row][col][8 channels]   (16 B per voxel per chunk).
// In the UMMA no-swizzle K-major canonical layout a row of the A operand is 16 B, 8 consecutive rows are contiguous
// (= 8 consecutive w), consecutive 8-row groups are SBO apart (= the window row pitch) and k-chunks are LBO apart
// (= the chunk-plane pitch).  Every tap is then just a different START ADDRESS into the same staged window: no
// im2col, no per-tap copies, no padding waste -- each input voxel is loaded once per tile (plus halo) with 16-byte
// LDGSTS (zero-filled outside the image, which implements the conv padding) and read by the tensor core straight
// from shared memory for every tap.
//   stride 1      : window 18 x (WT+2), tap (kd,kh,kw) -> plane d+kd-1, offset (kh, kw).
//   stride 2      : the producer de-interleaves the 33 x (2WT+1) window into 4 parity sub-planes, so that the rows
//                   2*ho+kh-1 of consecutive outputs are again consecutive: tap -> sub-plane (kh&1, kw&1), offset
//                   (kh>>1, kw>>1), plane 2*do+kd-1.
//   transposed s2 : GEMM rows are INPUT voxels q; the 8 output parity classes (rd,rh,rw) are 8 sub-convolutions with
//                   1/2/2/4/2/4/4/8 live taps (27 in total): out[2q+r] += in[q + (r==1 && k==0)] * W[k], k in
//                   {1} (r=0) or {0,2} (r=1).  One work item = one output plane (rd fixed), 4 (rh,rw) accumulators.
//
// Pipeline: 4 producer warps (cp.async) -> ring of NS plane slots (full/empty mbarriers) -> MMA warp (one elected
// lane issues tcgen05.mma, all operands warp-uniform; accumulators double-buffered in TMEM) -> 4 epilogue warps
// (tcgen05.ld, scale/shift, residual, ReLU, bf16 pack, 128-bit stores).  Persistent CTAs, one per SM; the weights of
// all taps stay resident in shared memory.
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"
#include <cstdlib>

namespace {

using namespace dpf;

constexpr int kEpiWarps = 4;
constexpr int kProdWarps = 4;
constexpr int kMmaWarp = kEpiWarps;                                   // warp 4
// MMA issue, template parameter NISS (number of issuing warps):
//   NISS = 1 (default): ONE elected thread issues every MMA of a work item in program order.  MMAs of one thread execute in
//            issue order, so the fp32 accumulation order -- and with it every output bit -- is fixed: run-to-run deterministic.
//   NISS = 4 (DPF_CONV_ISSUERS=4): the taps of a work item are dealt round-robin to four issuing warps (one warp walking 27 taps
//            -- tap decode + descriptor arithmetic + issue -- was the critical path of the stride-2 / transposed / 1x3x3 layers in
//            round 1).  Warps race into the same TMEM accumulator, so the accumulation ORDER varies from run to run: measured
//            on a B200 (tests/test_gpu_configs.py), ~1e-4 of the outputs differ by one bf16 ulp between two launches.
// In both modes no MMA overwrites: the epilogue leaves every accumulator zeroed after reading it (tcgen05.st).
constexpr int kGMmaWarps = 4;                                         // warps reserved for the role (NISS of them issue)
// Transposed kind: the epilogue was the bottleneck (ncu r01: LSU 64 %, tensor 13.6 % -- each of the 128 epilogue threads drains 4
// parity classes x 32 channels + 8 residual chunks per work item), so a SECOND set of four epilogue warps (warps 12-15, same TMEM
// lane quarters) takes the rh = 1 classes; the other geometries leave those warps idle.
constexpr int kEpi2Warps = 4;
constexpr int kEpi2First = kEpiWarps + kGMmaWarps + kProdWarps;      // warp 12
constexpr int kThreads = (kEpiWarps + kGMmaWarps + kProdWarps + kEpi2Warps) * 32;  // 512
constexpr int kMaxTaps = 27;

enum { GEO_S1 = 0, GEO_S2 = 1, GEO_T2 = 2 };

struct Tap {
  signed char dd;        // input plane relative to the work item's base plane
  unsigned char cls;     // accumulator class (transposed: rh*2+rw; else 0)
  unsigned short aoff;   // start offset inside a slot chunk-plane, in 16-byte units
  unsigned short widx;   // weight tile index (first tap slot of a fused group)
  unsigned short ncls;   // accumulator classes fed by ONE MMA (N = ncls * NPAD): transposed kind only, else 1
};

struct ConvKParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
  void* y;
  const float* scale;
  const float* shift;
  const void* residual;
  int B, D, H, W;                 // input grid
  int Do, Ho, Wo;                 // output grid
  int Mh, Mw, items;              // GEMM-row grid (h, w) and work items (planes) per tile
  int x_cstride, x_coff;          // channel stride / offset of the input tensor
  int cout, y_f32, y_cstride, y_coff, relu, res_pre;
  int nw;                         // number of weight tiles resident
  int tiles_h, tiles_w, ntiles;
  int debug;                      // DPF_CONV_DEBUG bits: 1 skip loads, 2 skip MMAs, 4 skip stores (timing experiments)
  // ---- kd-fused kernel only (defaults = dense 3-D tensor): strided views and the row-streamed 2-D mode -------------
  long long xs_b, xs_d, xs_h;     // voxel strides of the input: batch, plane, row (a row is W-contiguous)
  long long ys_b, ys_d, ys_h;     // voxel strides of the output and of the residual
  int halo;                       // 1: planes -1 and D exist (rows of a neighbouring stream) and are read; 0: they are padding
  int center_row_only;            // 1: the rows of a plane are independent streams -> only the kh = 1 in-plane taps
  int lin_d, lin_h, lin_max;      // lin_max > 0: (plane pl, row h) exists only if 0 <= pl*lin_d + h*lin_h < lin_max
  float slope;                    // activation when relu != 0: v > 0 ? v : slope * v   (0 = ReLU)
  int ntaps[2];
  Tap taps[2][kMaxTaps];          // program per work-item parity (only the transposed kind uses program 1)
};

template <int GEO, int CIN, int NPAD, int WT, int NS>
struct Cfg {
  static constexpr int NCH = CIN / 8;
  static constexpr int RWIN = (GEO == GEO_S1) ? 18 : (GEO == GEO_S2 ? 33 : 17);          // window rows loaded
  static constexpr int CWIN = (GEO == GEO_S1) ? WT + 2 : (GEO == GEO_S2 ? 2 * WT + 1 : WT + 1);
  static constexpr int NSUB = (GEO == GEO_S2) ? 4 : 1;
  static constexpr int RS = (GEO == GEO_S1) ? 18 : 17;                                     // rows per sub-plane
  static constexpr int WPS = (GEO == GEO_S1) ? WT + 2 : WT + 1;                            // row pitch (positions)
  static constexpr int SUB_POS = RS * WPS;
  static constexpr int PLANE_BYTES = NSUB * SUB_POS * 16;
  // chunk-plane pitch: 16-B multiple whose residue mod 128 spreads a quarter-warp's cp.async writes over all banks
  static constexpr int WANT = (NCH == 4) ? 32 : 16;
  static constexpr int CH_STRIDE = PLANE_BYTES + ((WANT - (PLANE_BYTES % 128)) + 128) % 128;
  static constexpr int SLOT_BYTES = NCH * CH_STRIDE;
  static constexpr int W_TAP_BYTES = NCH * NPAD * 16;
  static constexpr int W_BYTES = kMaxTaps * W_TAP_BYTES;
  static constexpr int NBLK = WT / 8;
  static constexpr int NCLS = (GEO == GEO_T2) ? 4 : 1;
  static constexpr int ACC_COLS = NCLS * NBLK * NPAD;
  static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
  static constexpr int KSTEPS = CIN / 16;
  static constexpr int LO_OFF = (GEO == GEO_T2) ? 0 : -1;       // lowest input plane of a work item, relative to its base
  static constexpr int SMEM_BYTES = W_BYTES + NS * SLOT_BYTES + 2 * NPAD * 4 + (2 * NS + 4) * 8 + 16 + 128;
  static_assert(2 * ACC_COLS <= 512, "accumulators do not fit TMEM");
  static_assert(CH_STRIDE % 16 == 0, "chunk pitch must be a 16-byte multiple");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
  static_assert(SLOT_BYTES / 16 < 65536, "tap offsets are 16-bit");
};

// Base input plane and tap program of a work item (= one output plane of the tile), per geometry.
template <int GEO>
__device__ __forceinline__ void item_planes(int item, int& base, int& prog) {
  if (GEO == GEO_S1) { base = item; prog = 0; }
  else if (GEO == GEO_S2) { base = 2 * item; prog = 0; }
  else { base = item >> 1; prog = item & 1; }
}

template <int GEO, int CIN, int NPAD, int WT, int NS, int NISS>
__global__ void __launch_bounds__(kThreads, 1) conv3d_tc_kernel(const __grid_constant__ ConvKParams p) {
  using C = Cfg<GEO, CIN, NPAD, WT, NS>;
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
    const int nbytes = p.nw * C::W_TAP_BYTES;
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
      mbar_init(&bar_empty[i], NISS);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_tfull[i], NISS);
      mbar_init(&bar_tempty[i], GEO == GEO_T2 ? kEpiWarps + kEpi2Warps : kEpiWarps);
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
  if (warp < kEpiWarps) {                                        // all accumulators start at zero (every MMA accumulates)
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c0 = 0; c0 < C::TMEM_COLS; c0 += 16) tmem_zero16(lane_base + c0);
    tmem_st_wait();
    tc_fence_before_sync();
  }
  __syncthreads();
  tc_fence_after_sync();

  const int D = p.D, H = p.H, W = p.W;

  if (warp >= kEpi2First && GEO != GEO_T2) {
    // second epilogue set: only the transposed kind uses it
  } else if (warp >= kMmaWarp + kGMmaWarps && warp < kEpi2First) {
    // =================================== producers: global -> shared ring ===================================
    const int pwarp = warp - (kMmaWarp + kGMmaWarps);            // 0..kProdWarps-1
    constexpr int PIECES_PER_ROW = C::CWIN * C::NCH;
    uint32_t g = 0;
    int prev_slot = -1;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int b = tile / (p.tiles_w * p.tiles_h);
      // window origin in input coordinates
      const int h0 = (GEO == GEO_S1) ? th * 16 - 1 : (GEO == GEO_S2 ? th * 32 - 1 : th * 16);
      const int w0 = (GEO == GEO_S1) ? tw * WT - 1 : (GEO == GEO_S2 ? tw * 2 * WT - 1 : tw * WT);
      for (int pl = 0; pl < D; ++pl, ++g) {
        const int slot = g % NS;
        const uint32_t ph = (g / NS) & 1u;
        mbar_wait(&bar_empty[slot], ph ^ 1u);
        const uint32_t sbase = smem_u32(s_slots + slot * C::SLOT_BYTES);
        const __nv_bfloat16* xplane = p.x + (static_cast<size_t>(b) * D + pl) * H * static_cast<size_t>(W) * p.x_cstride + p.x_coff;
        // row-based copy: one warp per window row (row decode once per row), lanes over the (column, channel-chunk) pieces
        if (!(p.debug & 1)) {
          for (int row = pwarp; row < C::RWIN; row += kProdWarps) {
            const int h = h0 + row;
            const bool hok = (h >= 0) && (h < H);
            const __nv_bfloat16* xrow = xplane + static_cast<size_t>(hok ? h : 0) * W * p.x_cstride;
            const int rowpos = (GEO == GEO_S2) ? (row & 1) * 2 * C::SUB_POS + (row >> 1) * C::WPS : row * C::WPS;
#pragma unroll
            for (int q = lane; q < PIECES_PER_ROW; q += 32) {
              const int col = q / C::NCH, c8 = q % C::NCH;      // NCH is 4 or 8: shifts
              const int w = w0 + col;
              const bool ok = hok && (w >= 0) && (w < W);
              const __nv_bfloat16* src = ok ? (xrow + static_cast<size_t>(w) * p.x_cstride + c8 * 8) : p.x;
              const int pos = rowpos + ((GEO == GEO_S2) ? (col & 1) * C::SUB_POS + (col >> 1) : col);
              cp_async16_zfill(sbase + c8 * C::CH_STRIDE + pos * 16, src, ok);
            }
          }
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
  } else if (warp >= kMmaWarp + NISS && warp < kMmaWarp + kGMmaWarps) {
    // reserved MMA-role warps that do not issue in this instantiation: nothing to do until the teardown barrier
  } else if (warp >= kMmaWarp && warp < kMmaWarp + kGMmaWarps) {
    // ============ MMA issuers: each warp runs the (warp-uniform) control flow of its taps, one elected lane issues =====
    const int mw = warp - kMmaWarp;
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, NPAD);
    const uint32_t wbase = smem_u32(s_w);
    const uint32_t sbase0 = smem_u32(s_slots);
    // descriptor templates: everything but the 14-bit start address
    const uint64_t adesc_hi = umma_desc_nosw(0, C::CH_STRIDE, C::WPS * 16);
    const uint64_t bdesc_hi = umma_desc_nosw(0, NPAD * 16, 128);
    const bool leader = elect_one();
    uint32_t g_base = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int nblk = min(C::NBLK, (p.Mw - tw * WT + 7) >> 3);
      int waited = -1;
      for (int item = 0; item < p.items; ++item, ++it) {
        const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
        mbar_wait(&bar_tempty[as], aph ^ 1u);
        tc_fence_after_sync();
        int base, prog;
        item_planes<GEO>(item, base, prog);
        const int nt = p.ntaps[prog];
        for (int t = mw; t < nt; t += NISS) {
          const Tap tp = p.taps[prog][t];
          const int pin = base + tp.dd;
          if (pin < 0 || pin >= D) continue;
          const uint32_t gp = g_base + pin;
          const uint32_t slot = gp % NS;
          if (pin > waited) {
            mbar_wait(&bar_full[slot], (gp / NS) & 1u);
            tc_fence_after_sync();
            waited = pin;
          }
          const uint32_t a0 = ((sbase0 + slot * C::SLOT_BYTES) >> 4) + tp.aoff;
          const uint32_t b0 = (wbase + tp.widx * C::W_TAP_BYTES) >> 4;
          const uint32_t acc0 = tmem_base + (as * C::NCLS + tp.cls) * C::NBLK * NPAD;
          if (leader && !(p.debug & 2)) {
            if (GEO == GEO_T2 && C::NBLK == 1) {
              // transposed kind: the taps of the (up to 4) parity classes that read the SAME input shift are one MMA with
              // N = ncls * NPAD -- their accumulators are adjacent TMEM columns and their weight tiles one [Cin x N] operand --
              // so the A window is read 15 instead of 27 times per pair of output planes
              const uint32_t ng = tp.ncls * NPAD;
              const uint32_t idg = umma_idesc_bf16_f32(128, static_cast<int>(ng));
              const uint64_t bhi = umma_desc_nosw(0, ng * 16, 128);
#pragma unroll
              for (int ks = 0; ks < C::KSTEPS; ++ks) {
                const uint64_t bdesc = bhi | static_cast<uint64_t>(b0 + ks * 2 * ng);
                const uint64_t adesc = adesc_hi | static_cast<uint64_t>(a0 + ks * 2 * (C::CH_STRIDE >> 4));
                umma_bf16(acc0, adesc, bdesc, idg, true);
              }
            } else {
#pragma unroll
            for (int ks = 0; ks < C::KSTEPS; ++ks) {
              // (shared-memory addresses >> 4 are < 2^14: the start-address field cannot overflow, no masking needed)
              const uint64_t bdesc = bdesc_hi | static_cast<uint64_t>(b0 + ks * 2 * NPAD);
#pragma unroll
              for (int blk = 0; blk < C::NBLK; ++blk) {
                if (blk == 0 || blk < nblk) {                   // a tile always has its first block (no runtime test)
                  const uint64_t adesc = adesc_hi | static_cast<uint64_t>(a0 + ks * 2 * (C::CH_STRIDE >> 4) + blk * 8);
                  umma_bf16(acc0 + blk * NPAD, adesc, bdesc, idesc, true);
                }
              }
            }
            }
          }
        }
        // release the input planes that no later work item of this tile needs
        const int lo_cur = max(base + C::LO_OFF, 0);
        int lo_next = D;
        if (item + 1 < p.items) {
          int nb, np;
          item_planes<GEO>(item + 1, nb, np);
          lo_next = min(max(nb + C::LO_OFF, 0), D);
        }
        if (leader) {
          umma_commit(&bar_tfull[as]);
          for (int q = lo_cur; q < lo_next; ++q) umma_commit(&bar_empty[(g_base + q) % NS]);
        }
        __syncwarp();
      }
      g_base += D;
    }
  } else {
    // =================================== epilogue: TMEM -> registers -> global ============================
    // warps 0-3: all classes (transposed kind: the rh = 0 classes); warps 12-15 (transposed kind only): the rh = 1 classes
    uint32_t it = 0;
    const int ew = warp & 3;                                     // TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31)
    const int cls_lo = (GEO == GEO_T2 && warp >= kEpi2First) ? 2 : 0;
    const int cls_hi = (GEO == GEO_T2 && warp < kEpi2First) ? 2 : C::NCLS;
    const int m = ew * 32 + lane;
    const int hrow = m >> 3, wcol = m & 7;
    constexpr int OS = (GEO == GEO_T2) ? 2 : 1;                  // output stride of the GEMM-row grid
    constexpr bool kPrefetch = (GEO == GEO_T2) && (C::NBLK == 1) && (NPAD == 32);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int b = tile / (p.tiles_w * p.tiles_h);
      const int nblk = min(C::NBLK, (p.Mw - tw * WT + 7) >> 3);
      const int mh = th * 16 + hrow;
      for (int item = 0; item < p.items; ++item, ++it) {
        const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
        // Transposed kind: a lane's 8 residual chunks (4 parity classes x 2 channel chunks) do not depend on the MMAs, so they
        // are requested BEFORE waiting for the accumulators; fetching them one by one inside the loop below (8 dependent
        // ~600-cycle round trips per work item on only four epilogue warps) was longer than the item's MMA phase.
        uint4 rr[kPrefetch ? 8 : 1][2];
        const bool use_pref = kPrefetch && p.residual != nullptr && !p.y_f32 && p.cout == NPAD && (((p.y_cstride | p.y_coff) & 15) == 0);
        if (kPrefetch && use_pref) {
#pragma unroll
          for (int cls = 0; cls < C::NCLS; ++cls) {
            if (cls < cls_lo || cls >= cls_hi) continue;
            const int oh = mh * OS + (cls >> 1), mw = tw * WT + wcol, ow = mw * OS + (cls & 1);
            const bool ok = (mh < p.Mh) && (mw < p.Mw) && (oh < p.Ho) && (ow < p.Wo);
            const size_t vox = ((static_cast<size_t>(b) * p.Do + item) * p.Ho + oh) * static_cast<size_t>(p.Wo) + ow;
            const __nv_bfloat16* ro = reinterpret_cast<const __nv_bfloat16*>(p.residual) + vox * p.y_cstride + p.y_coff;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              rr[(cls * 2 + q) % (kPrefetch ? 8 : 1)][0] = make_uint4(0u, 0u, 0u, 0u);
              rr[(cls * 2 + q) % (kPrefetch ? 8 : 1)][1] = make_uint4(0u, 0u, 0u, 0u);
              if (ok) ld_global_v8(ro + q * 16, rr[(cls * 2 + q) % (kPrefetch ? 8 : 1)][0], rr[(cls * 2 + q) % (kPrefetch ? 8 : 1)][1]);
            }
          }
        }
        mbar_wait(&bar_tfull[as], aph);
        tc_fence_after_sync();
#pragma unroll(kPrefetch ? 4 : 1)
        for (int cls = 0; cls < C::NCLS; ++cls) {
          if (cls < cls_lo || cls >= cls_hi) continue;
          const int oh = mh * OS + (cls >> 1);
#pragma unroll 1
          for (int blk = 0; blk < nblk; ++blk) {
            const int mw = tw * WT + blk * 8 + wcol;
            const int ow = mw * OS + (cls & 1);
            const bool ok = (mh < p.Mh) && (mw < p.Mw) && (oh < p.Ho) && (ow < p.Wo);
            const size_t vox = ((static_cast<size_t>(b) * p.Do + item) * p.Ho + oh) * static_cast<size_t>(p.Wo) + ow;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + ((as * C::NCLS + cls) * C::NBLK + blk) * NPAD;
#pragma unroll
            for (int c0 = 0; c0 < NPAD; c0 += 16) {
              uint32_t v[16];
              __syncwarp();                                      // tcgen05.ld is .sync.aligned: keep the warp converged
              tmem_ld16(taddr + c0, v);
              tmem_ld_wait();
              tmem_zero16(taddr + c0);                           // leave the accumulator zeroed for its next work item
              if (!ok || c0 >= p.cout || (p.debug & 4)) continue;
              const int n = min(16, p.cout - c0);
              float f[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
              if (p.y_f32) {
                float* yo = reinterpret_cast<float*>(p.y) + vox * p.y_cstride + p.y_coff + c0;
                const float* ro = p.residual ? reinterpret_cast<const float*>(p.residual) + vox * p.y_cstride + p.y_coff + c0 : nullptr;
                if (n == 16 && ((p.y_cstride | p.y_coff) & 3) == 0) {   // 16-byte aligned full chunk: 128-bit loads / stores
#pragma unroll
                  for (int j4 = 0; j4 < 16; j4 += 4) {
                    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ro) r = *reinterpret_cast<const float4*>(ro + j4);
                    float o[4] = {f[j4], f[j4 + 1], f[j4 + 2], f[j4 + 3]};
                    const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      if (ro && p.res_pre) o[k] += rr[k];
                      o[k] = o[k] * s_scale[c0 + j4 + k] + s_shift[c0 + j4 + k];
                      if (ro && !p.res_pre) o[k] += rr[k];
                      if (p.relu) o[k] = fmaxf(o[k], 0.f);
                    }
                    *reinterpret_cast<float4*>(yo + j4) = make_float4(o[0], o[1], o[2], o[3]);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {                  // fully unrolled + predicated: keeps f[] in registers
                    if (j < n) {
                      float val = f[j];
                      if (ro && p.res_pre) val += ro[j];
                      val = val * s_scale[c0 + j] + s_shift[c0 + j];
                      if (ro && !p.res_pre) val += ro[j];
                      if (p.relu) val = fmaxf(val, 0.f);
                      yo[j] = val;
                    }
                  }
                }
              } else {
                __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(p.y) + vox * p.y_cstride + p.y_coff + c0;
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = f[j] * s_scale[c0 + j] + s_shift[c0 + j];
                const bool wide = (n == 16) && (((p.y_cstride | p.y_coff) & 15) == 0);   // 32-byte aligned full chunk
                if (p.residual) {
                  const __nv_bfloat16* ro = reinterpret_cast<const __nv_bfloat16*>(p.residual) + vox * p.y_cstride + p.y_coff + c0;
                  uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;
                  if (kPrefetch && use_pref) {
                    r0 = rr[(cls * 2 + (c0 >> 4)) % (kPrefetch ? 8 : 1)][0];
                    r1 = rr[(cls * 2 + (c0 >> 4)) % (kPrefetch ? 8 : 1)][1];
                  } else if (wide) ld_global_v8(ro, r0, r1);
                  else {
                    r0 = *reinterpret_cast<const uint4*>(ro);
                    if (n > 8) r1 = *reinterpret_cast<const uint4*>(ro + 8);
                  }
                  f[0] += bf16_lo(r0.x); f[1] += bf16_hi(r0.x); f[2] += bf16_lo(r0.y); f[3] += bf16_hi(r0.y);
                  f[4] += bf16_lo(r0.z); f[5] += bf16_hi(r0.z); f[6] += bf16_lo(r0.w); f[7] += bf16_hi(r0.w);
                  f[8] += bf16_lo(r1.x); f[9] += bf16_hi(r1.x); f[10] += bf16_lo(r1.y); f[11] += bf16_hi(r1.y);
                  f[12] += bf16_lo(r1.z); f[13] += bf16_hi(r1.z); f[14] += bf16_lo(r1.w); f[15] += bf16_hi(r1.w);
                }
                if (p.relu) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                }
                uint4 o0, o1;                                      // n is a multiple of 8 (checked on the host)
                o0.x = pack_bf16x2(f[0], f[1]); o0.y = pack_bf16x2(f[2], f[3]); o0.z = pack_bf16x2(f[4], f[5]); o0.w = pack_bf16x2(f[6], f[7]);
                o1.x = pack_bf16x2(f[8], f[9]); o1.y = pack_bf16x2(f[10], f[11]); o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
                if (wide) st_global_v8(yo, o0, o1);
                else {
                  *reinterpret_cast<uint4*>(yo) = o0;
                  if (n > 8) *reinterpret_cast<uint4*>(yo + 8) = o1;
                }
              }
            }
          }
        }
        tmem_st_wait();
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

// ============================================================================================================
// kd-fused variant for the 3x3x3 stride-1 layers (the bulk of the FLOPs, all with small Cout).
//
// With N = Cout = 32 the A operand dominates shared-memory traffic: a 128x32x16 MMA reads 4 KB of A for 1 KB of B.
// The three depth taps kd = 0,1,2 of one (kh,kw) read the SAME A window of input plane p and contribute to output
// planes p+1, p, p-1.  Keeping the accumulators of consecutive output planes in CONSECUTIVE TMEM column groups
// (ring of R stages, stage(Q) = (-Q) mod R) lets ONE tcgen05.mma with N = 3*Cout and B = [W(kd=0) | W(kd=1) | W(kd=2)]
// do all three at once: A traffic / 3, 9 x Cin/16 MMAs per input plane instead of 27 x.  A side effect: every input plane
// is consumed in a single pass, so its slot is released immediately (shallow ring).  Since one instruction now
// mixes "first contribution" and "accumulate" columns, accumulators are zeroed by the epilogue (tcgen05.st) when it
// drains them, and every MMA accumulates.  When the three stages wrap around the ring the MMA is split in two.
// ============================================================================================================
constexpr int kFEpiWarps = 8;                                       // fused kernel: 8 epilogue warps, MMA warps 8-9, 4 producers
constexpr int kFMmaWarp = kFEpiWarps;
// MMA issue (template parameter NISS, see the generic kernel): 1 = one thread, program order, deterministic;
// NISS == NBLK (default where a tile has 2 or 3 blocks): one issuing warp per 128-row block, each accumulator owned by ONE thread --
// as deterministic as a single issuer, with 2-3 warps sharing the descriptor arithmetic;
// 4 (DPF_CONV_ISSUERS=4) = the (128-row block, K-step) pairs of a plane are dealt round-robin to four issuing warps -- K-steps of one block then race
// into one accumulator (every MMA accumulates into a pre-zeroed stage, so any order is CORRECT, but the fp32 rounding differs from
// run to run).  Round 1 used 4 because the single issuer was the critical path (~10 uniform-datapath instructions of descriptor
// arithmetic per tcgen05.mma behind a per-MMA branch); the single-issuer loop now has no branch and no address masking, so the
// compiler can software-pipeline the descriptor arithmetic of consecutive MMAs.
constexpr int kFMmaWarps = 4;                                       // warps reserved for the role (NISS of them issue)
constexpr int kFThreads = (kFEpiWarps + kFMmaWarps + kProdWarps) * 32;   // 512

template <int CIN, int NPAD, int WT, int NS, int R, int DILW = 1, int KS = 1>
struct FCfg {
  static constexpr int NCH = CIN / 8;
  static constexpr int WP = WT + 2 * DILW;                       // window columns: DILW = column dilation (2-D mode only)
  static constexpr int PLANE_BYTES = 18 * WP * 16;
  static constexpr int WANT = (NCH == 4) ? 32 : 16;
  static constexpr int CH_STRIDE = PLANE_BYTES + ((WANT - (PLANE_BYTES % 128)) + 128) % 128;
  static constexpr int SLOT_BYTES = NCH * CH_STRIDE;
  static constexpr int W_ROWS = 3 * NPAD;                        // B rows per (kh,kw): [kd][co]
  static constexpr int W_GROUP_BYTES = NCH * W_ROWS * 16;        // one (kh,kw): [c8][kd*NPAD+co][8]
  static constexpr int W_BYTES = 9 * W_GROUP_BYTES;
  static constexpr int NBLK = WT / 8;
  static constexpr int SET_COLS = R * NBLK * NPAD;               // one accumulator ring
  static constexpr int COLS = KS * SET_COLS;                     // KS = 2: split-K, one ring per issuing warp (summed in the epilogue)
  static constexpr int TMEM_COLS = COLS <= 32 ? 32 : COLS <= 64 ? 64 : COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
  static constexpr int KSTEPS = CIN / 16;
  static constexpr int SMEM_BYTES = W_BYTES + NS * SLOT_BYTES + 2 * NPAD * 4 + (2 * NS + 2 * R) * 8 + 16 + 128;
  static_assert(COLS <= 512, "accumulator ring does not fit TMEM");
  static_assert(R >= 4, "need >= 4 accumulator stages");
  static_assert(3 * NPAD <= 256, "fused N too large");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

template <int CIN, int NPAD, int WT, int NS, int R, int DILW, int NISS>
__global__ void __launch_bounds__(kFThreads, 1) conv3d_kdfused_kernel(const __grid_constant__ ConvKParams p) {
  // One-block tiles (WT = 8) with two issuing warps: SPLIT-K -- warp 0 issues the first half of the K-steps into accumulator ring 0,
  // warp 1 the second half into ring 1, the epilogue adds the two rings in a fixed order.  Every accumulator is still written by
  // exactly one thread in program order (deterministic), and two warps share the descriptor arithmetic.
  constexpr int KS = (NISS == 2 && WT == 8) ? 2 : 1;
  using C = FCfg<CIN, NPAD, WT, NS, R, DILW, KS>;
  static_assert(KS == 1 || C::KSTEPS % 2 == 0, "split-K needs an even number of K-steps");
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

  // ---- one-time setup: weights global [tap=(kd,kh,kw)][c8][NPAD][8] -> smem [khkw][c8][kd*NPAD+n][8] --------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    constexpr int PER_TAP = C::NCH * NPAD;                       // 16-byte rows per tap
    for (int i = threadIdx.x; i < 27 * PER_TAP; i += kFThreads) {
      const int tap = i / PER_TAP, rem = i - tap * PER_TAP;
      const int c8 = rem / NPAD, n = rem - c8 * NPAD;
      const int kd = tap / 9, khkw = tap - kd * 9;
      dst[(khkw * C::NCH + c8) * C::W_ROWS + kd * NPAD + n] = __ldg(src + i);
    }
    for (int i = threadIdx.x; i < NPAD; i += kFThreads) {
      s_scale[i] = (p.scale != nullptr && i < p.cout) ? p.scale[i] : 1.0f;
      s_shift[i] = (p.shift != nullptr && i < p.cout) ? p.shift[i] : 0.0f;
    }
    fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&bar_full[i], kProdWarps);
      mbar_init(&bar_empty[i], NISS);
    }
    for (int i = 0; i < R; ++i) {
      mbar_init(&bar_tfull[i], NISS);
      mbar_init(&bar_tempty[i], kFEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == kFMmaWarp) {
    tmem_alloc(s_tmem, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const int D = p.D, H = p.H, W = p.W;

  if (warp >= kFMmaWarp + kFMmaWarps) {
    // =================================== producers (identical to the generic kernel) ======================
    const int pwarp = warp - (kFMmaWarp + kFMmaWarps);
    constexpr int PIECES_PER_ROW = C::WP * C::NCH;
    uint32_t g = 0;
    int prev_slot = -1;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int b = tile / (p.tiles_w * p.tiles_h);
      const int h0 = th * 16 - 1, w0 = tw * WT - DILW;
      for (int pl = -p.halo; pl < D + p.halo; ++pl, ++g) {
        const int slot = g % NS;
        mbar_wait(&bar_empty[slot], ((g / NS) & 1u) ^ 1u);
        const uint32_t sbase = smem_u32(s_slots + slot * C::SLOT_BYTES);
        const long long plane_off = static_cast<long long>(b) * p.xs_b + static_cast<long long>(pl) * p.xs_d;
        if (!(p.debug & 1)) {                                    // row-based copy, see the generic kernel
          for (int row = pwarp; row < 18; row += kProdWarps) {
            const int h = h0 + row;
            const bool hok = (h >= 0) && (h < H) &&
                             (p.lin_max == 0 || static_cast<unsigned>(pl * p.lin_d + h * p.lin_h) < static_cast<unsigned>(p.lin_max));
            const __nv_bfloat16* xrow = p.x + (hok ? (plane_off + static_cast<long long>(h) * p.xs_h) * p.x_cstride : 0) + p.x_coff;
#pragma unroll
            for (int q = lane; q < PIECES_PER_ROW; q += 32) {
              const int col = q / C::NCH, c8 = q % C::NCH;
              const int w = w0 + col;
              const bool ok = hok && (w >= 0) && (w < W);
              const __nv_bfloat16* src = ok ? (xrow + static_cast<size_t>(w) * p.x_cstride + c8 * 8) : p.x;
              cp_async16_zfill(sbase + c8 * C::CH_STRIDE + (row * C::WP + col) * 16, src, ok);
            }
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
    if (prev_slot >= 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[prev_slot]);
    }
  } else if (warp >= kFMmaWarp + NISS) {
    // reserved MMA-role warps that do not issue in this instantiation
  } else if (warp >= kFMmaWarp) {
    // =================================== MMA issuers (NISS warps, disjoint shares of every plane) =========
    const int mw = warp - kFMmaWarp;
    const uint32_t wbase = smem_u32(s_w) >> 4;
    const uint32_t sbase0 = smem_u32(s_slots);
    const uint64_t adesc_hi = umma_desc_nosw(0, C::CH_STRIDE, C::WP * 16);
    const uint64_t bdesc_hi = umma_desc_nosw(0, C::W_ROWS * 16, 128);
    const bool leader = elect_one();
    // Input plane pl feeds output planes pl+1-kd (kd = 0,1,2) that lie in [0, D).  Input planes are counted for the slot ring
    // (g_in), output planes for the accumulator ring (g_out); with halo planes (row-streamed 2-D mode) pl runs from -1 to D.
    uint32_t g_in = 0, g_out = 0;
    const int HALO = p.halo;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int nblk = min(C::NBLK, (p.Mw - tw * WT + 7) >> 3);
      for (int pl = -HALO; pl < D + HALO; ++pl, ++g_in) {
        const uint32_t slot = g_in % NS;
        // accumulator stages touched for the first time by this input plane must have been drained (and zeroed)
        if (HALO == 0 && pl == 0) mbar_wait(&bar_tempty[(R - g_out % R) % R], (g_out / R) & 1u);
        if (pl + 1 < D) {
          const uint32_t Qn = g_out + static_cast<uint32_t>(pl + 1);
          mbar_wait(&bar_tempty[(R - Qn % R) % R], (Qn / R) & 1u);
        }
        mbar_wait(&bar_full[slot], (g_in / NS) & 1u);
        tc_fence_after_sync();
        const int kd_lo = (pl + 1 < D) ? 0 : pl + 2 - D, kd_hi = min(2, pl + 1);
        const int n_kd = kd_hi - kd_lo + 1;
        const uint32_t Qf = g_out + static_cast<uint32_t>(pl + 1 - kd_lo);
        const uint32_t s_first = (R - Qf % R) % R;                   // stage of output plane pl+1-kd_lo; next kd -> next stage
        const int run1 = min(n_kd, static_cast<int>(R - s_first)), run2 = n_kd - run1;
        const uint32_t idesc1 = umma_idesc_bf16_f32(128, run1 * NPAD);
        const uint32_t idesc2 = umma_idesc_bf16_f32(128, (run2 > 0 ? run2 : 1) * NPAD);
        const uint32_t a_slot = (sbase0 + slot * C::SLOT_BYTES) >> 4;
        if (leader && !(p.debug & 2)) {
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh) {
            if (p.center_row_only && kh != 1) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const uint32_t a0 = a_slot + kh * C::WP + kw * DILW;
              const uint32_t b0 = wbase + (kh * 3 + kw) * (C::W_GROUP_BYTES >> 4) + kd_lo * NPAD;
#pragma unroll
              for (int ks = 0; ks < C::KSTEPS; ++ks) {
                const uint64_t bd1 = bdesc_hi | static_cast<uint64_t>(b0 + ks * 2 * C::W_ROWS);
                const uint64_t bd2 = bdesc_hi | static_cast<uint64_t>(b0 + ks * 2 * C::W_ROWS + run1 * NPAD);
#pragma unroll
                for (int blk = 0; blk < C::NBLK; ++blk) {
                  // NISS == NBLK: warp mw OWNS 128-row block mw (all of its K-steps and taps) -- every accumulator is written by one
                  // thread in program order: parallel issue AND deterministic.  Otherwise (NISS = 4, opt-in) round-robin shares.
                  if (KS == 2 ? (ks / (C::KSTEPS / 2) != mw)
                              : (NISS > 1 && (NISS == C::NBLK ? (blk != mw) : ((blk * C::KSTEPS + ks) % NISS != mw)))) continue;
                  if (blk == 0 || blk < nblk) {                   // a tile always has its first block (no runtime test)
                    const uint64_t adesc = adesc_hi | static_cast<uint64_t>(a0 + ks * 2 * (C::CH_STRIDE >> 4) + blk * 8);
                    const uint32_t col = tmem_base + (KS == 2 ? mw * C::SET_COLS : 0) + blk * (R * NPAD);
                    umma_bf16(col + s_first * NPAD, adesc, bd1, idesc1, true);
                    if (run2 > 0) umma_bf16(col, adesc, bd2, idesc2, true);
                  }
                }
              }
            }
          }
        }
        if (leader) {
          umma_commit(&bar_empty[slot]);                               // this input plane is fully consumed
          if (pl >= 1) {                                               // output plane pl-1 complete
            const uint32_t Qd = g_out + static_cast<uint32_t>(pl - 1);
            umma_commit(&bar_tfull[(R - Qd % R) % R]);
          }
          if (HALO == 0 && pl == D - 1) {                              // no plane D: the last output plane is complete too
            const uint32_t Qd = g_out + static_cast<uint32_t>(pl);
            umma_commit(&bar_tfull[(R - Qd % R) % R]);
          }
        }
        __syncwarp();
      }
      g_out += D;
    }
  } else {
    // =================================== epilogue (8 warps: 2 per TMEM lane quarter) =======================
    // A warp may only touch TMEM lanes 32*(warp%4)..+31; the two warps of a quarter split the (block, 16-column chunk)
    // items of an output plane.  All of a warp's TMEM loads are issued before the single wait.
    constexpr int CHUNKS = NPAD / 16;
    constexpr int ITEMS = C::NBLK * CHUNKS;
    constexpr int PER = (ITEMS + 1) / 2;
    const int quarter = warp & 3, half = warp >> 2;
    const int m = quarter * 32 + lane;
    const int hrow = m >> 3, wcol = m & 7;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    // zero the accumulator ring once (each half its share of the columns), then publish every stage as free
    for (int c0 = half * 16; c0 < C::COLS; c0 += 32) tmem_zero16(lane_base + c0);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0)
      for (int s = 0; s < R; ++s) mbar_arrive(&bar_tempty[s]);
    uint32_t g_base = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int b = tile / (p.tiles_w * p.tiles_h);
      const int nblk = min(C::NBLK, (p.Mw - tw * WT + 7) >> 3);
      const int h = th * 16 + hrow;
      for (int d = 0; d < D; ++d) {
        const uint32_t Q = g_base + d;
        const uint32_t st = (R - Q % R) % R;
        // The residual of this output plane does not depend on the MMAs: request it BEFORE waiting for the accumulator (bf16 output,
        // full 32-byte chunks).  Fetched inside the store loop it was one dependent global round trip per item on the epilogue warps
        // (dres1.2: 0.296 ms against 0.209 ms for the same layer without a residual).
        const bool res_pref = (p.residual != nullptr) && !p.y_f32 && (((p.y_cstride | p.y_coff) & 15) == 0) && (p.cout % 16 == 0);
        uint4 rr[PER][2];
        if (res_pref) {
#pragma unroll
          for (int i = 0; i < PER; ++i) {
            const int item = i * 2 + half;
            const int blk = item / CHUNKS, c0 = (item % CHUNKS) * 16;
            const int w = tw * WT + blk * 8 + wcol;
            const bool ok = (item < ITEMS) && (blk < nblk) && (h < H) && (w < W) && (c0 < p.cout) &&
                            (p.lin_max == 0 || static_cast<unsigned>(d * p.lin_d + h * p.lin_h) < static_cast<unsigned>(p.lin_max));
            rr[i][0] = make_uint4(0u, 0u, 0u, 0u);
            rr[i][1] = rr[i][0];
            if (ok) {
              const size_t vox = static_cast<size_t>(static_cast<long long>(b) * p.ys_b + static_cast<long long>(d) * p.ys_d +
                                                     static_cast<long long>(h) * p.ys_h + w);
              ld_global_v8(reinterpret_cast<const __nv_bfloat16*>(p.residual) + vox * p.y_cstride + p.y_coff + c0, rr[i][0], rr[i][1]);
            }
          }
        }
        mbar_wait(&bar_tfull[st], (Q / R) & 1u);
        tc_fence_after_sync();
        uint32_t v[PER][16];
        uint32_t v2[KS == 2 ? PER : 1][16];                          // split-K: the second accumulator ring
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          const int item = i * 2 + half;
          const int blk = item / CHUNKS, c0 = (item % CHUNKS) * 16;
          if (item < ITEMS && blk < nblk) {
            tmem_ld16(lane_base + blk * (R * NPAD) + st * NPAD + c0, v[i]);
            if (KS == 2) tmem_ld16(lane_base + C::SET_COLS + blk * (R * NPAD) + st * NPAD + c0, v2[i % (KS == 2 ? PER : 1)]);
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < PER; ++i) {                              // leave the stage zeroed for its next output plane
          const int item = i * 2 + half;
          const int blk = item / CHUNKS, c0 = (item % CHUNKS) * 16;
          if (item < ITEMS && blk < nblk) {
            tmem_zero16(lane_base + blk * (R * NPAD) + st * NPAD + c0);
            if (KS == 2) tmem_zero16(lane_base + C::SET_COLS + blk * (R * NPAD) + st * NPAD + c0);
          }
        }
        if (KS == 2) {                                               // ring 0 + ring 1, always in this order
#pragma unroll
          for (int i = 0; i < PER; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j)
              v[i][j] = __float_as_uint(__uint_as_float(v[i][j]) + __uint_as_float(v2[i % (KS == 2 ? PER : 1)][j]));
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          const int item = i * 2 + half;
          const int blk = item / CHUNKS, c0 = (item % CHUNKS) * 16;
          const int w = tw * WT + blk * 8 + wcol;
          const bool ok = (item < ITEMS) && (blk < nblk) && (h < H) && (w < W) &&
                          (p.lin_max == 0 || static_cast<unsigned>(d * p.lin_d + h * p.lin_h) < static_cast<unsigned>(p.lin_max));
          if (!ok || c0 >= p.cout || (p.debug & 4)) continue;
          const size_t vox = static_cast<size_t>(static_cast<long long>(b) * p.ys_b + static_cast<long long>(d) * p.ys_d +
                                                 static_cast<long long>(h) * p.ys_h + w);
          const int n = min(16, p.cout - c0);
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[i][j]);
          if (p.y_f32) {
            float* yo = reinterpret_cast<float*>(p.y) + vox * p.y_cstride + p.y_coff + c0;
            const float* ro = p.residual ? reinterpret_cast<const float*>(p.residual) + vox * p.y_cstride + p.y_coff + c0 : nullptr;
            if (n == 16 && ((p.y_cstride | p.y_coff) & 3) == 0) {     // 16-byte aligned full chunk: 128-bit loads / stores
#pragma unroll
              for (int j4 = 0; j4 < 16; j4 += 4) {
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ro) r = *reinterpret_cast<const float4*>(ro + j4);
                float o[4] = {f[j4], f[j4 + 1], f[j4 + 2], f[j4 + 3]};
                const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (ro && p.res_pre) o[k] += rr[k];
                  o[k] = o[k] * s_scale[c0 + j4 + k] + s_shift[c0 + j4 + k];
                  if (ro && !p.res_pre) o[k] += rr[k];
                  if (p.relu) o[k] = fmaxf(o[k], 0.f);
                }
                *reinterpret_cast<float4*>(yo + j4) = make_float4(o[0], o[1], o[2], o[3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {                      // fully unrolled + predicated: keeps f[] in registers
                if (j < n) {
                  float val = f[j];
                  if (ro && p.res_pre) val += ro[j];
                  val = val * s_scale[c0 + j] + s_shift[c0 + j];
                  if (ro && !p.res_pre) val += ro[j];
                  if (p.relu) val = fmaxf(val, 0.f);
                  yo[j] = val;
                }
              }
            }
          } else {
            __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(p.y) + vox * p.y_cstride + p.y_coff + c0;
            if (p.scale != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = f[j] * s_scale[c0 + j] + s_shift[c0 + j];
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] += s_shift[c0 + j];
            }
            const bool wide = (n == 16) && (((p.y_cstride | p.y_coff) & 15) == 0);   // 32-byte aligned full chunk
            if (p.residual) {
              const __nv_bfloat16* ro = reinterpret_cast<const __nv_bfloat16*>(p.residual) + vox * p.y_cstride + p.y_coff + c0;
              uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;
              if (res_pref) { r0 = rr[i][0]; r1 = rr[i][1]; }
              else if (wide) ld_global_v8(ro, r0, r1);
              else {
                r0 = *reinterpret_cast<const uint4*>(ro);
                if (n > 8) r1 = *reinterpret_cast<const uint4*>(ro + 8);
              }
              f[0] += bf16_lo(r0.x); f[1] += bf16_hi(r0.x); f[2] += bf16_lo(r0.y); f[3] += bf16_hi(r0.y);
              f[4] += bf16_lo(r0.z); f[5] += bf16_hi(r0.z); f[6] += bf16_lo(r0.w); f[7] += bf16_hi(r0.w);
              f[8] += bf16_lo(r1.x); f[9] += bf16_hi(r1.x); f[10] += bf16_lo(r1.y); f[11] += bf16_hi(r1.y);
              f[12] += bf16_lo(r1.z); f[13] += bf16_hi(r1.z); f[14] += bf16_lo(r1.w); f[15] += bf16_hi(r1.w);
            }
            if (p.relu) {
              const float sl = p.slope;                              // 0 = ReLU, else LeakyReLU / single-parameter PReLU
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f) + sl * fminf(f[j], 0.f);
            }
            uint4 o0, o1;
            o0.x = pack_bf16x2(f[0], f[1]); o0.y = pack_bf16x2(f[2], f[3]); o0.z = pack_bf16x2(f[4], f[5]); o0.w = pack_bf16x2(f[6], f[7]);
            o1.x = pack_bf16x2(f[8], f[9]); o1.y = pack_bf16x2(f[10], f[11]); o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
            if (wide) st_global_v8(yo, o0, o1);
            else {
              *reinterpret_cast<uint4*>(yo) = o0;
              if (n > 8) *reinterpret_cast<uint4*>(yo + 8) = o1;
            }
          }
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[st]);
      }
      g_base += D;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kFMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// DPF_CONV_ISSUERS = 1 (default: deterministic single issuer) | 4 (round-1 four-warp issue, run-to-run rounding differences)
int conv_issuers() {
  static int n = 0;
  if (n == 0) { const char* e = getenv("DPF_CONV_ISSUERS"); n = (e && atoi(e) == 4) ? 4 : 1; }
  return n;
}

template <int CIN, int NPAD, int WT, int NS, int R, int DILW, int NISS>
int launch_fused_n(ConvKParams kp, cudaStream_t st) {
  using C = FCfg<CIN, NPAD, WT, NS, R, DILW, (NISS == 2 && WT == 8) ? 2 : 1>;
  kp.tiles_h = (kp.Mh + 15) / 16;
  kp.tiles_w = (kp.Mw + WT - 1) / WT;
  kp.ntiles = kp.B * kp.tiles_h * kp.tiles_w;
  auto kern = conv3d_kdfused_kernel<CIN, NPAD, WT, NS, R, DILW, NISS>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return dpf::fail("dpf_conv3d_fwd: cannot opt in to %d B shared memory: %s", C::SMEM_BYTES, cudaGetErrorString(e));
    attr_done = true;
  }
  const int grid = std::min(kp.ntiles, dpf::sm_count());
  kern<<<grid, kFThreads, C::SMEM_BYTES, st>>>(kp);
  return dpf::after_launch("dpf_conv3d_fwd");
}

template <int CIN, int NPAD, int WT, int NS, int R, int DILW = 1>
int launch_fused(const ConvKParams& kp, cudaStream_t st) {
  constexpr int NBLK = WT / 8;
  // one issuing warp per block when a tile has several blocks; one-block tiles with >= 4 K-steps (Cin = 64): two warps, split-K
  constexpr int OWN = (NBLK >= 2 && NBLK <= 4) ? NBLK : ((CIN == 64 && 2 * R * NPAD <= 512) ? 2 : 1);
  return conv_issuers() == 4 ? launch_fused_n<CIN, NPAD, WT, NS, R, DILW, 4>(kp, st) : launch_fused_n<CIN, NPAD, WT, NS, R, DILW, OWN>(kp, st);
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
template <int GEO, int CIN, int NPAD, int WT, int NS, int NISS>
int launch_n(ConvKParams kp, cudaStream_t st) {
  using C = Cfg<GEO, CIN, NPAD, WT, NS>;
  // tap start offsets inside a slot chunk-plane (16-byte units); the builders stash (dh << 8 | dw) in aoff
  for (int pr = 0; pr < 2; ++pr)
    for (int t = 0; t < kp.ntaps[pr]; ++t) {
      Tap& tp = kp.taps[pr][t];
      const int dh = tp.aoff >> 8, dw = tp.aoff & 0xFF;
      int off;
      if (GEO == GEO_S2) off = ((dh & 1) * 2 + (dw & 1)) * C::SUB_POS + (dh >> 1) * C::WPS + (dw >> 1);
      else off = dh * C::WPS + dw;
      tp.aoff = static_cast<unsigned short>(off);
    }
  kp.tiles_h = (kp.Mh + 15) / 16;
  kp.tiles_w = (kp.Mw + WT - 1) / WT;
  kp.ntiles = kp.B * kp.tiles_h * kp.tiles_w;
  auto kern = conv3d_tc_kernel<GEO, CIN, NPAD, WT, NS, NISS>;
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

template <int GEO, int CIN, int NPAD, int WT, int NS>
int launch(const ConvKParams& kp, cudaStream_t st) {
  return conv_issuers() == 4 ? launch_n<GEO, CIN, NPAD, WT, NS, 4>(kp, st) : launch_n<GEO, CIN, NPAD, WT, NS, 1>(kp, st);
}

int npad_for(int cout) { return cout <= 16 ? 16 : (cout <= 32 ? 32 : 64); }

inline Tap mk_tap(int dd, int cls, int dh, int dw, int widx) {
  Tap t;
  t.dd = static_cast<signed char>(dd);
  t.cls = static_cast<unsigned char>(cls);
  t.aoff = static_cast<unsigned short>((dh << 8) | dw);
  t.widx = static_cast<unsigned short>(widx);
  t.ncls = 1;
  return t;
}

}  // namespace

extern "C" long long dpf_conv3d_weight_elems(int kind, int Cin, int Cout) {
  const int ntaps = (kind == 0 || kind == 1 || kind == 2) ? 27 : (kind == 3 ? 9 : 1);
  return static_cast<long long>(ntaps) * Cin * npad_for(Cout);
}

extern "C" int dpf_conv3d_fwd(const dpf_conv3d_args* a, void* stream) {
  DPF_REQUIRE(a != nullptr, "dpf_conv3d_fwd: null args");
  DPF_REQUIRE(a->x && a->w && a->y, "dpf_conv3d_fwd: null tensor pointer");
  DPF_REQUIRE(DPF_ALIGNED16(a->x) && DPF_ALIGNED16(a->w) && DPF_ALIGNED16(a->y), "dpf_conv3d_fwd: pointers must be 16-byte aligned");
  DPF_REQUIRE(a->kind >= 0 && a->kind <= 4, "dpf_conv3d_fwd: kind %d not in 0..4", a->kind);
  DPF_REQUIRE(a->Cin == 32 || a->Cin == 64, "dpf_conv3d_fwd: Cin=%d must be 32 or 64 (per launch)", a->Cin);
  DPF_REQUIRE(a->Cout >= 1 && a->Cout <= 64, "dpf_conv3d_fwd: Cout=%d must be in [1,64] (split wider layers on the host)", a->Cout);
  DPF_REQUIRE(a->B > 0 && a->D > 0 && a->D <= 64 && a->H > 0 && a->W > 0, "dpf_conv3d_fwd: bad shape");
  DPF_REQUIRE(a->y_f32 || (a->Cout % 8 == 0 && a->y_cstride % 8 == 0 && a->y_coff % 8 == 0),
              "dpf_conv3d_fwd: bf16 output needs Cout, y_cstride, y_coff multiples of 8");
  DPF_REQUIRE(a->y_cstride >= a->y_coff + a->Cout, "dpf_conv3d_fwd: y_cstride too small");
  const int x_cstride = a->x_cstride > 0 ? a->x_cstride : a->Cin;
  DPF_REQUIRE(x_cstride % 8 == 0 && a->x_coff % 8 == 0 && a->x_coff + a->Cin <= x_cstride, "dpf_conv3d_fwd: bad input channel window");
  DPF_REQUIRE(a->stats == nullptr, "dpf_conv3d_fwd: fused batch statistics are not built yet");
  ConvKParams kp{};
  kp.x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  kp.w = reinterpret_cast<const __nv_bfloat16*>(a->w);
  kp.y = a->y;
  kp.scale = a->scale;
  kp.shift = a->shift;
  kp.residual = a->residual;
  kp.B = a->B; kp.D = a->D; kp.H = a->H; kp.W = a->W;
  kp.x_cstride = x_cstride; kp.x_coff = a->x_coff;
  kp.cout = a->Cout; kp.y_f32 = a->y_f32; kp.y_cstride = a->y_cstride; kp.y_coff = a->y_coff; kp.relu = a->relu;
  kp.res_pre = a->res_pre;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("DPF_CONV_DEBUG"); dbg = e ? atoi(e) : 0; }
    kp.debug = dbg;
  }
  int geo = GEO_S1;
  int t = 0;
  if (a->kind == 0 || a->kind == 3 || a->kind == 4) {
    kp.Do = a->D; kp.Ho = a->H; kp.Wo = a->W;
    if (a->kind == 0) {
      for (int kd = 0; kd < 3; ++kd)
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw, ++t) kp.taps[0][t] = mk_tap(kd - 1, 0, kh, kw, t);
    } else if (a->kind == 3) {
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw, ++t) kp.taps[0][t] = mk_tap(0, 0, kh, kw, t);
    } else {
      kp.taps[0][t] = mk_tap(0, 0, 1, 1, 0);
      t = 1;
    }
    kp.ntaps[0] = t; kp.nw = t;
    kp.Mh = kp.Ho; kp.Mw = kp.Wo; kp.items = kp.Do;
  } else if (a->kind == 1) {
    geo = GEO_S2;
    kp.Do = (a->D + 1) / 2; kp.Ho = (a->H + 1) / 2; kp.Wo = (a->W + 1) / 2;
    for (int kd = 0; kd < 3; ++kd)
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw, ++t) kp.taps[0][t] = mk_tap(kd - 1, 0, kh, kw, t);
    kp.ntaps[0] = 27; kp.nw = 27;
    kp.Mh = kp.Ho; kp.Mw = kp.Wo; kp.items = kp.Do;
  } else {
    geo = GEO_T2;
    kp.Do = 2 * a->D; kp.Ho = 2 * a->H; kp.Wo = 2 * a->W;
    // out[2q+r] += in[q + o] * W[k]:  r=0 -> (k=1,o=0);  r=1 -> (k=0,o=1), (k=2,o=0)
    const int ks[2][2] = {{1, -1}, {0, 2}};
    const int os[2][2] = {{0, 0}, {1, 0}};
    const int nk[2] = {1, 2};
    // Fused groups: for every (output-plane parity rd, depth tap id) and every in-plane input shift (oh, ow), the parity classes
    // (rh, rw) with rh in R(oh), rw in R(ow), R(0) = {0, 1}, R(1) = {1}, read the same A window; runs of ADJACENT classes
    // (cls = 2 rh + rw) are one MMA.  Group order = weight-slot order of the packed tensor (dpf_conv3d_t2_weight_order below /
    // ops.fuse_t2_weight): (0,0) -> classes 0..3; (0,1) -> {1}, {3}; (1,0) -> {2, 3}; (1,1) -> {3}.
    {
      int slot = 0;
      for (int rd = 0; rd < 2; ++rd) {
        int n = 0;
        for (int id = 0; id < nk[rd]; ++id) {
          const int runs[5][4] = {{0, 0, 0, 4}, {0, 1, 1, 1}, {0, 1, 3, 1}, {1, 0, 2, 2}, {1, 1, 3, 1}};   // oh, ow, first class, ncls
          for (int r = 0; r < 5; ++r, ++n) {
            Tap tp = mk_tap(os[rd][id], runs[r][2], runs[r][0], runs[r][1], slot);
            tp.ncls = static_cast<unsigned short>(runs[r][3]);
            kp.taps[rd][n] = tp;
            slot += runs[r][3];
          }
        }
        kp.ntaps[rd] = n;
      }
    }
    kp.nw = 27;
    kp.Mh = a->H; kp.Mw = a->W; kp.items = kp.Do;
  }
  kp.xs_h = a->W; kp.xs_d = static_cast<long long>(a->H) * a->W; kp.xs_b = kp.xs_d * a->D;   // dense [B,D,H,W] views
  kp.ys_h = kp.xs_h; kp.ys_d = kp.xs_d; kp.ys_b = kp.xs_b;
  kp.halo = 0; kp.center_row_only = 0; kp.lin_d = kp.lin_h = kp.lin_max = 0; kp.slope = a->slope;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int npad = npad_for(a->Cout);
  DPF_REQUIRE(a->slope == 0.f || (a->kind == 0 && npad <= 32 && a->relu),
              "dpf_conv3d_fwd: a LeakyReLU slope is built for the kd-fused 3x3x3 stride-1 layers (Cout <= 32) only");
  if (geo == GEO_S1) {
    DPF_REQUIRE(a->Cin == 32 || a->Cout <= 32, "dpf_conv3d_fwd: Cin=64 supports Cout<=32 per launch (split on the host)");
    static int fused = -1;
    if (fused < 0) { const char* e = getenv("DPF_CONV_FUSED"); fused = e ? atoi(e) : 1; }
    if (a->kind == 0 && fused && npad <= 32) {                     // kd-fused issue for the 3x3x3 stride-1 layers
      // 32 -> 32: 24-wide tiles (3 blocks, 5-stage accumulator ring) when they do not pad the width more than 16-wide ones:
      // 22 % instead of 27 % halo traffic and fewer, larger tiles per CTA (1036 vs 1006 TFLOP/s at W = 420)
      const bool wide24 = ((a->W + 23) / 24) * 24 <= ((a->W + 15) / 16) * 16;
      if (a->Cin == 32 && npad == 32 && wide24) return launch_fused<32, 32, 24, 4, 5>(kp, st);
      if (a->Cin == 32 && npad == 32) return launch_fused<32, 32, 16, 4, 8>(kp, st);
      if (a->Cin == 32 && npad == 16 && wide24) return launch_fused<32, 16, 24, 4, 8>(kp, st);
      if (a->Cin == 32 && npad == 16) return launch_fused<32, 16, 16, 4, 8>(kp, st);
      // (64 -> 32 with 16-wide tiles -- 2 blocks, one issuing warp each -- only fits a 2-slot ring next to the 110 KB of weights:
      //  measured 714 vs 949 TFLOP/s for the 8-wide tile with a 4-slot ring, so it stays 8-wide / single issuer)
      if (a->Cin == 64 && npad == 32) return launch_fused<64, 32, 8, 4, 8>(kp, st);
      if (a->Cin == 64 && npad == 16) return launch_fused<64, 16, 8, 4, 16>(kp, st);
    }
    if (a->Cin == 32 && npad == 32) return launch<GEO_S1, 32, 32, 24, 5>(kp, st);
    if (a->Cin == 32 && npad == 16) return launch<GEO_S1, 32, 16, 24, 5>(kp, st);
    if (a->Cin == 32 && npad == 64) return launch<GEO_S1, 32, 64, 8, 6>(kp, st);
    if (a->Cin == 64 && npad == 32) return launch<GEO_S1, 64, 32, 8, 5>(kp, st);
    if (a->Cin == 64 && npad == 16) return launch<GEO_S1, 64, 16, 8, 5>(kp, st);
  } else if (geo == GEO_S2) {
    DPF_REQUIRE(a->Cin == 32 && a->Cout <= 32, "dpf_conv3d_fwd: stride-2 kind supports Cin=32, Cout<=32 per launch (split on the host)");
    if (npad == 32) return launch<GEO_S2, 32, 32, 8, 4>(kp, st);
    if (npad == 16) return launch<GEO_S2, 32, 16, 8, 4>(kp, st);
  } else {
    DPF_REQUIRE(a->Cout <= 32, "dpf_conv3d_fwd: transposed kind supports Cout<=32 per launch (split on the host)");
    if (a->Cin == 64 && npad == 32) return launch<GEO_T2, 64, 32, 8, 4>(kp, st);
    if (a->Cin == 32 && npad == 32) return launch<GEO_T2, 32, 32, 8, 4>(kp, st);
  }
  return dpf::fail("dpf_conv3d_fwd: no kernel for kind=%d Cin=%d Cout=%d", a->kind, a->Cin, a->Cout);
}

// 2-D 3x3 convolution (stride 1, pad = dilation) on channels-last images, on the kd-fused kernel: the image is read as 16
// independent row streams (H' = 16 segments of L consecutive rows, stream position = "depth"), so that a GEMM block is 16 streams
// x 8 pixels, the image's kh taps are the fused depth taps (N = 3*Cout, accumulator ring over output rows) and only the centre
// in-plane row taps are issued.  Planes -1 and L of a stream are the neighbouring streams' rows (halo planes).
// Dilation d: the rows of one residue class (row mod d) form a dilation-1 problem of their own (plane stride d*W), so the kh
// dilation costs d launches over 1/d of the rows each; the kw dilation is a template parameter of the window / tap offsets.
extern "C" int dpf_conv2d_fwd(const void* x, const void* w, void* y, const float* scale, const float* shift, const void* residual,
                              int N, int H, int W, int Cin, int Cout, int x_cstride, int x_coff, int y_cstride, int y_coff,
                              int dil, int relu, float slope, void* stream) {
  DPF_REQUIRE(x && w && y, "dpf_conv2d_fwd: null tensor pointer");
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(w) && DPF_ALIGNED16(y), "dpf_conv2d_fwd: pointers must be 16-byte aligned");
  DPF_REQUIRE((Cin == 32 || Cin == 64) && Cout >= 8 && Cout <= 32 && Cout % 8 == 0,
              "dpf_conv2d_fwd: Cin=%d Cout=%d (built: Cin 32 | 64, Cout <= 32 per launch, multiple of 8)", Cin, Cout);
  DPF_REQUIRE(dil == 1 || ((dil == 3 || dil == 5) && Cin == 32), "dpf_conv2d_fwd: dilation %d (built: 1; 3 and 5 for Cin = 32)", dil);
  DPF_REQUIRE(N > 0 && H > 0 && W > 0, "dpf_conv2d_fwd: bad shape");
  DPF_REQUIRE(x_cstride % 8 == 0 && x_coff % 8 == 0 && x_coff + Cin <= x_cstride && y_cstride % 8 == 0 && y_coff % 8 == 0 &&
              y_coff + Cout <= y_cstride, "dpf_conv2d_fwd: bad channel windows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int npad = npad_for(Cout);
  const bool wide24 = ((W + 23) / 24) * 24 <= ((W + 15) / 16) * 16;
  for (int res = 0; res < dil && res < H; ++res) {
    const int Hres = (H - res + dil - 1) / dil;                 // rows of this residue class
    const int L = (Hres + 15) / 16;
    ConvKParams kp{};
    const size_t row0 = static_cast<size_t>(res) * W;           // first voxel of the class
    kp.x = reinterpret_cast<const __nv_bfloat16*>(x) + row0 * x_cstride;
    kp.w = reinterpret_cast<const __nv_bfloat16*>(w);
    kp.y = reinterpret_cast<__nv_bfloat16*>(y) + row0 * y_cstride;
    kp.scale = scale; kp.shift = shift;
    kp.residual = residual ? reinterpret_cast<const __nv_bfloat16*>(residual) + row0 * y_cstride : nullptr;
    kp.B = N; kp.D = L; kp.H = 16; kp.W = W;
    kp.Do = L; kp.Ho = 16; kp.Wo = W; kp.Mh = 16; kp.Mw = W; kp.items = L;
    kp.x_cstride = x_cstride; kp.x_coff = x_coff;
    kp.cout = Cout; kp.y_f32 = 0; kp.y_cstride = y_cstride; kp.y_coff = y_coff; kp.relu = relu; kp.res_pre = 0; kp.nw = 27;
    kp.xs_d = static_cast<long long>(dil) * W; kp.xs_h = kp.xs_d * L; kp.xs_b = static_cast<long long>(H) * W;
    kp.ys_h = kp.xs_h; kp.ys_d = kp.xs_d; kp.ys_b = kp.xs_b;
    kp.halo = 1; kp.center_row_only = 1; kp.lin_d = 1; kp.lin_h = L; kp.lin_max = Hres; kp.slope = slope;
    { const char* e = getenv("DPF_CONV_DEBUG"); kp.debug = e ? atoi(e) : 0; }
    int rc;
    if (dil == 3) rc = npad == 32 ? launch_fused<32, 32, 24, 4, 5, 3>(kp, st) : launch_fused<32, 16, 24, 4, 8, 3>(kp, st);
    else if (dil == 5) rc = npad == 32 ? launch_fused<32, 32, 24, 4, 5, 5>(kp, st) : launch_fused<32, 16, 24, 4, 8, 5>(kp, st);
    else if (Cin == 64) rc = npad == 32 ? launch_fused<64, 32, 8, 4, 8>(kp, st) : launch_fused<64, 16, 8, 4, 16>(kp, st);
    else if (npad == 32) rc = wide24 ? launch_fused<32, 32, 24, 4, 5>(kp, st) : launch_fused<32, 32, 16, 4, 8>(kp, st);
    else rc = wide24 ? launch_fused<32, 16, 24, 4, 8>(kp, st) : launch_fused<32, 16, 16, 4, 8>(kp, st);
    if (rc) return rc;
  }
  return 0;
}
