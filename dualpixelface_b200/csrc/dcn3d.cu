// 3-D deformable convolution (D3D) forward for sm_100a: trilinear gather fused into the A-operand producer of a
// tcgen05 GEMM -- no [27*Cin, B*D*H*W] column buffer (7-13 GB fp32 in the reference at 1120x1680).
//
// Replaces DCN.deform_conv_forward (src/module/dcn3d/src/deform_conv.h:10-29 -> src/cuda/deform_conv_cuda.cu:18-126)
// and its im2col kernel (src/cuda/deform_im2col_cuda.cuh:192-265, sampling rule :26-72, in-bounds test :248).
//   y[v, o] = relu?( scale[o] * sum_{tap,c} W[o,c,tap] * trilinear(x[:, c], p_v + tap - 1 + offset[v, 3*tap + (0,1,2)]) + shift[o] )
// x [B,D,H,W,x_cstride] bf16 (the first CINP channels are gathered; zero-padded beyond the real Cin), offset [B,D,H,W,off_cstride >= 81] fp32 ((d,h,w) per tap), y [B,D,H,W,64] bf16.
//
// Work unit: a 16 x 16 spatial tile of one depth plane = 256 voxels (2 GEMM blocks of 128 rows); every CTA owns a
// contiguous range of units so that the planes a tap gathers from stay hot in L1 / L2 across taps and units.  For every tap, 16 producer warps compute the
// trilinear sample of all CINP channels (4 threads per voxel, 32 bytes each; packed-bf16 blend) and write it
// as the bf16 A tile in the UMMA no-swizzle K-major layout; the tap's weight tile [CINP x 64] is streamed next to it by a
// 1-D TMA bulk copy.  One elected lane issues the tcgen05.mma; accumulators are double-buffered in TMEM so the
// epilogue of a unit overlaps the gather of the next.
//
// Offsets, STAGED variant (DPF_DCN_STAGED=1; off_cstride a multiple of 4 and >= 84) -- built, tested, measured, and OFF by default.
// A quarter of the kernel's L1 wavefronts are the OFFSET loads: 3 scalar LDGs per (voxel, tap) in which the 4 lanes of a voxel read
// the same word, i.e. 8 lines = 8 wavefronts per warp instruction for 32 useful bytes (profiles/r02_dcn3d_source.txt).  The staged
// variant keeps the unit's 256 x 81 offsets in shared memory (16-byte cp.async in two halves, taps 0-11 | 12-26, each half
// re-filled for the NEXT unit while the other is in use; voxel pitches of 36 / 52 words = conflict-free), which does cut the global
// load wavefronts by 20 % (234 M -> 187 M per launch) -- but the extra 88 KB of shared memory come out of the L1 cache the gather
// lives on: L1 hit rate 83 % -> 43 %, L2->L1 sectors x3.1, and the launch is not faster alone (3.51 vs 3.40 ms) and 30 % slower inside
// the model, where the offsets are larger (profiles/r02_dcn3d_staged.txt).  The gather needs its L1 more than its LSU slots.
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"
#include <cstdlib>

namespace {

using namespace dpf;

constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kProdWarps = 16;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;   // 672
#ifndef DPF_DCN_GROUPS
#define DPF_DCN_GROUPS 2
#endif
constexpr int kGroups = DPF_DCN_GROUPS;                       // producer groups: group i gathers the taps t = i (mod kGroups)
static_assert(kGroups == 1 || kGroups == 2, "one or two producer groups (every group needs a pipeline stage of its own in flight)");
constexpr int kNOut = 64;
#ifndef DPF_DCN_BLOCKS
#define DPF_DCN_BLOCKS 1
#endif
constexpr int kBlocks = DPF_DCN_BLOCKS;                       // GEMM blocks (128 voxels = 8 rows x 16 columns each) per work unit
constexpr int kUnitH = 8 * kBlocks;                           // rows of a work unit (its width is 16)
constexpr int kUnitVox = 128 * kBlocks;
#ifndef DPF_DCN_STAGES
#define DPF_DCN_STAGES 3
#endif
constexpr int kStages = DPF_DCN_STAGES;
static_assert(kStages >= kGroups + 1, "every producer group needs a stage of its own in flight next to the one the MMA reads");
constexpr int kTaps = 27;
// staged offsets: half A = taps 0..11 (36 floats per voxel), half B = taps 12..26 (45 floats, copied as 48 and padded to 52)
constexpr int kSplitTap = 12;
constexpr int kOffPitchA = 36, kOffPitchB = 52;              // words; 36 mod 32 = 4, 52 mod 32 = 20: 8 consecutive voxels -> 8 banks
constexpr int kOffPiecesA = 9, kOffPiecesB = 12;             // 16-byte pieces per voxel
constexpr int kOffBytes = kUnitVox * (kOffPitchA + kOffPitchB) * 4;

// (kd, kh, kw) - 1 of tap t = (kd*3 + kh)*3 + kw
__constant__ signed char kTapD[27] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1};
__constant__ signed char kTapH[27] = {-1, -1, -1, 0, 0, 0, 1, 1, 1, -1, -1, -1, 0, 0, 0, 1, 1, 1, -1, -1, -1, 0, 0, 0, 1, 1, 1};
__constant__ signed char kTapW[27] = {-1, 0, 1, -1, 0, 1, -1, 0, 1, -1, 0, 1, -1, 0, 1, -1, 0, 1, -1, 0, 1, -1, 0, 1, -1, 0, 1};

struct DcnParams {
  const __nv_bfloat16* x;
  const float* offset;
  const __nv_bfloat16* w;
  const float* scale;
  const float* shift;
  __nv_bfloat16* y;
  int B, D, H, W, relu, x_cstride, off_cstride;
  long long nvox;
  int nunits, tiles_h, tiles_w;
  int skip_oob;
};

template <int CINP, int OMODE = 0>
struct DCfg {
  static constexpr bool STAGED = OMODE == 1;
  static constexpr int NCH = CINP / 8;
  static constexpr int KSTEPS = CINP / 16;
  static constexpr int A_CHUNK_BYTES = 128 * 16 + 16;                // chunk pitch (= LBO), +16 B: conflict-free st.shared
  static constexpr int A_BLOCK_BYTES = NCH * A_CHUNK_BYTES;          // [chunk][128 rows][16 B]
  static constexpr int A_STAGE_BYTES = kBlocks * A_BLOCK_BYTES;
  static constexpr int W_TAP_BYTES = NCH * kNOut * 16;               // [chunk][64 rows][16 B]
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + W_TAP_BYTES;
  static constexpr int TMEM_COLS = 2 * kBlocks * kNOut;                // 2 stages x kBlocks blocks x 64 columns (128 or 256)
  static constexpr int BASE_BYTES = kStages * STAGE_BYTES + 2 * kNOut * 4 + (2 * kStages + 4) * 8 + 16;
  static constexpr int OFF_AT = (BASE_BYTES + 15) & ~15;             // staged offsets (16-byte aligned)
  static constexpr int SMEM_BYTES = (STAGED ? OFF_AT + kOffBytes : BASE_BYTES) + 128;
  static_assert(CINP % 16 == 0, "CINP must be a multiple of 16");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

// unit -> (d, h0, w0, b): depth fastest, so that consecutive units of a CTA share two of their three input planes
__device__ __forceinline__ void unit_coords(int unit, const DcnParams& p, int& d, int& h0, int& w0, int& b) {
  d = unit % p.D;
  int t = unit / p.D;
  w0 = (t % p.tiles_w) * 16;
  t /= p.tiles_w;
  h0 = (t % p.tiles_h) * kUnitH;
  b = t / p.tiles_h;
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// base + 32-bit byte offset as one IMAD.WIDE.U32 (the compiler's 64-bit add is two instructions per gathered corner)
__device__ __forceinline__ const char* ptr_add_u32(const char* base, uint32_t off) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(off), "l"(reinterpret_cast<unsigned long long>(base)));
  return reinterpret_cast<const char*>(r);
}

// producers only (16 warps): named barrier 1
__device__ __forceinline__ void prod_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kProdWarps * 32) : "memory"); }

// cp.async one half (pieces per voxel NP, first float F0, smem pitch PITCH words) of a unit's offsets
template <int NP, int F0, int PITCH>
__device__ __forceinline__ void stage_offsets(const DcnParams& p, int unit, int ptid, float* s_dst) {
  int ud, uh0, uw0, ub;
  {
    ud = unit % p.D;
    int t = unit / p.D;
    uw0 = (t % p.tiles_w) * 16;
    t /= p.tiles_w;
    uh0 = (t % p.tiles_h) * kUnitH;
    ub = t / p.tiles_h;
  }
  const long long plane = (static_cast<long long>(ub) * p.D + ud) * p.H;
  const uint32_t dst0 = smem_u32(s_dst);
  for (int i = ptid; i < kUnitVox * NP; i += kProdWarps * 32) {
    const int r = i / NP, q = i - r * NP;
    const int hh = uh0 + (r >> 4), ww = uw0 + (r & 15);
    const bool ok = (hh < p.H) && (ww < p.W);
    const float* src = ok ? p.offset + ((plane + hh) * p.W + ww) * p.off_cstride + F0 + 4 * q : p.offset;
    cp_async16_zfill(dst0 + (r * PITCH + 4 * q) * 4, src, ok);
  }
  cp_async_commit();
}

template <int CINP, int OMODE>
__global__ void __launch_bounds__(kThreads, 1) dcn3d_kernel(const __grid_constant__ DcnParams p) {
  using C = DCfg<CINP, OMODE>;
  constexpr bool STAGED = OMODE == 1;
  constexpr bool VECOFF = OMODE == 2;
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
  float* s_offA = reinterpret_cast<float*>(smem + C::OFF_AT);            // [256][kOffPitchA]   (STAGED only)
  float* s_offB = s_offA + kUnitVox * kOffPitchA;                             // [256][kOffPitchB]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < kNOut; i += kThreads) {
    s_scale[i] = p.scale ? p.scale[i] : 1.0f;
    s_shift[i] = p.shift ? p.shift[i] : 0.0f;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bar_full[i], kProdWarps / (STAGED ? 1 : kGroups) + 1);   // the gather warps of one group + the weight-copy issuer (expect_tx)
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
  // contiguous unit range per CTA (neighbouring tiles / depth planes back to back: L1 / L2 reuse of the gathered planes)
  const int per_cta = (p.nunits + gridDim.x - 1) / gridDim.x;
  const int unit_lo = min(static_cast<int>(blockIdx.x) * per_cta, p.nunits), unit_hi = min(unit_lo + per_cta, p.nunits);

  if (warp > kMmaWarp) {
    // ======================= producers: trilinear gather -> bf16 A tile; weight tile by TMA ==================
    // Producer groups: the 16 gather warps are split into NG groups; group i produces the taps t = i (mod NG) of every unit, each
    // of its threads covering PASSES = NG voxel rows per tap.  A stage then needs only the warps of ONE group (the per-tap coupling
    // of all 16 warps is gone: while one group waits for its gathered lines the other one blends), and the two voxel rows of a
    // thread are independent gathers the compiler can overlap.
    constexpr int NG = STAGED ? 1 : kGroups;
    constexpr int kVoxPerPass = kProdWarps * 32 / 4 / NG;    // 4 threads (32 bytes = two channel chunks each) per voxel
    const int ptid_all = threadIdx.x - (kMmaWarp + 1) * 32;  // 0..511
    const int grp = ptid_all / (kProdWarps * 32 / NG);
    const int ptid = ptid_all - grp * (kProdWarps * 32 / NG); // thread index inside the group
    const int cq = ptid & 3;                                  // which pair of 16-byte channel chunks
    const int vsub = ptid >> 2;
    // 4 lanes per voxel, 32 bytes (two 16-byte channel chunks) each: one LDG.256 per corner per lane, a voxel's 128-byte
    // line is touched exactly once per corner, and the per-(voxel,tap) coordinate / weight arithmetic is replicated
    // 4x instead of 8x.  All addresses are 32-bit byte offsets from the (uniform) tensor base.
    constexpr int PASSES = kBlocks * 128 / kVoxPerPass;
    static_assert(C::NCH % 2 == 0, "channel chunks are handled in pairs");
    const int HW = H * W;
    const uint32_t cs2 = static_cast<uint32_t>(p.x_cstride) * 2u;            // bytes per voxel
    const bool lane_live = (2 * cq) < C::NCH;
    const char* xbytes = reinterpret_cast<const char*>(p.x);
    const uint32_t lane_off = (lane_live ? cq : 0) * 32u;
    uint32_t gbase = 0;
    if (STAGED && unit_lo < unit_hi) {                       // both halves of the first unit
      stage_offsets<kOffPiecesA, 0, kOffPitchA>(p, unit_lo, ptid, s_offA);
      stage_offsets<kOffPiecesB, 3 * kSplitTap, kOffPitchB>(p, unit_lo, ptid, s_offB);
      cp_async_wait<1>();                                    // half A landed (half B may still be in flight)
      prod_barrier();
    }
    for (int unit = unit_lo; unit < unit_hi; ++unit) {
      // unit = 16 x 16 spatial tile of one depth plane; row r of the unit -> (h0 + r/16, w0 + r%16).  Voxel
      // coordinates of this thread's PASSES rows are decoded once per unit and reused by all 27 taps.
      int ud, uh0, uw0, ub;
      unit_coords(unit, p, ud, uh0, uw0, ub);
      int vh[PASSES], vw[PASSES];
      bool vlive[PASSES];
      const int vbase = ub * D * HW;                                  // voxel index of (b, 0, 0, 0); < 2^31 (host check)
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int r = ps * kVoxPerPass + vsub;
        const int hh = uh0 + (r >> 4), ww = uw0 + (r & 15);
        vlive[ps] = (hh < H) && (ww < W) && lane_live;
        vh[ps] = vlive[ps] ? hh : 0;
        vw[ps] = vlive[ps] ? ww : 0;
      }
      // the (d, h, w) offsets of tap t+1 are requested while tap t is gathered: offset load -> address -> gather would
      // otherwise be two dependent memory round trips per (voxel, tap)
      const float* obase[PASSES];
      float onext[PASSES][3];
      // VECOFF: the 81 offsets of a voxel are read as 16-byte pieces, 3 pieces (= 4 taps x 3 floats) at a time: lane cq < 3 of the
      // voxel's 4-lane group loads piece cq of the block (ONE load instruction = 8 lines per warp for 4 taps, instead of 3 scalar
      // loads = 24 lines per tap); the three floats of a tap are then fetched from the owning lane with warp shuffles.
      float4 oblk[PASSES], onxt[PASSES];
      const float4* ob4[PASSES];
      if (!STAGED) {
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
          const int r = ps * kVoxPerPass + vsub;
          const int hh = uh0 + (r >> 4), ww = uw0 + (r & 15);
          const bool vin = (hh < H) && (ww < W);                          // independent of lane_live (the lane may serve offsets only)
          obase[ps] = p.offset + static_cast<size_t>(vbase + (ud * H + (vin ? hh : 0)) * W + (vin ? ww : 0)) * p.off_cstride;
          if (VECOFF) {
            ob4[ps] = reinterpret_cast<const float4*>(obase[ps]) + min(cq, 2);
            oblk[ps] = __ldg(ob4[ps]);
          } else {
            onext[ps][0] = __ldg(obase[ps] + 3 * grp + 0); onext[ps][1] = __ldg(obase[ps] + 3 * grp + 1);
            onext[ps][2] = __ldg(obase[ps] + 3 * grp + 2);
          }
        }
      }
      for (int tap = grp; tap < kTaps; tap += NG) {
        const uint32_t g = gbase + tap;
        const int stage = g % kStages;
        const uint32_t ph = (g / kStages) & 1u;
        float ocur[PASSES][3];
        if (VECOFF) {
          const int j = tap & 3;
          if (j == grp && (tap - j) + 4 < kTaps) {                     // this group's first tap of a 4-tap block: request the next block
#pragma unroll
            for (int ps = 0; ps < PASSES; ++ps) onxt[ps] = __ldg(ob4[ps] + 3 * ((tap >> 2) + 1));
          }
          const int gl = lane & ~3;
          // word 3*j + a of the 12-word block lives in lane (3*j + a) / 4, component (3*j + a) % 4: warp-uniform switch on j
#define DPF_OFF_PICK(A, LANE, COMP) ocur[ps][A] = __shfl_sync(0xffffffffu, oblk[ps].COMP, gl | LANE)
#pragma unroll
          for (int ps = 0; ps < PASSES; ++ps) {
            switch (j) {
              case 0: DPF_OFF_PICK(0, 0, x); DPF_OFF_PICK(1, 0, y); DPF_OFF_PICK(2, 0, z); break;
              case 1: DPF_OFF_PICK(0, 0, w); DPF_OFF_PICK(1, 1, x); DPF_OFF_PICK(2, 1, y); break;
              case 2: DPF_OFF_PICK(0, 1, z); DPF_OFF_PICK(1, 1, w); DPF_OFF_PICK(2, 2, x); break;
              default: DPF_OFF_PICK(0, 2, y); DPF_OFF_PICK(1, 2, z); DPF_OFF_PICK(2, 2, w); break;
            }
          }
#undef DPF_OFF_PICK
          if (j == 4 - NG + grp) {                                     // this group's last tap of the block
#pragma unroll
            for (int ps = 0; ps < PASSES; ++ps) oblk[ps] = onxt[ps];
          }
        } else if (STAGED) {
          if (tap == kSplitTap) {
            // half B of this unit has landed and every producer is past its last read of half A: refill A for the next unit
            cp_async_wait<0>();
            prod_barrier();
            if (unit + 1 < unit_hi) stage_offsets<kOffPiecesA, 0, kOffPitchA>(p, unit + 1, ptid, s_offA);
          }
          const uint32_t so = (tap < kSplitTap) ? smem_u32(s_offA + 3 * tap) : smem_u32(s_offB + 3 * (tap - kSplitTap));
          const int pitch = (tap < kSplitTap) ? kOffPitchA : kOffPitchB;
#pragma unroll
          for (int ps = 0; ps < PASSES; ++ps) {
            const uint32_t o = so + static_cast<uint32_t>((ps * kVoxPerPass + vsub) * pitch) * 4u;
            ocur[ps][0] = lds_f32(o); ocur[ps][1] = lds_f32(o + 4); ocur[ps][2] = lds_f32(o + 8);
          }
        } else {
#pragma unroll
          for (int ps = 0; ps < PASSES; ++ps) {
            ocur[ps][0] = onext[ps][0]; ocur[ps][1] = onext[ps][1]; ocur[ps][2] = onext[ps][2];
            if (tap + NG < kTaps) {
              const float* on = obase[ps] + (tap + NG) * 3;
              onext[ps][0] = __ldg(on + 0); onext[ps][1] = __ldg(on + 1); onext[ps][2] = __ldg(on + 2);
            }
          }
        }
        mbar_wait(&bar_empty[stage], ph ^ 1u);
        uint8_t* sa = s_stage + stage * C::STAGE_BYTES;
        if (ptid == 0) {
          mbar_arrive_expect_tx(&bar_full[stage], C::W_TAP_BYTES);
          bulk_g2s(smem_u32(sa + C::A_STAGE_BYTES), p.w + static_cast<size_t>(tap) * (C::W_TAP_BYTES / 2), C::W_TAP_BYTES, &bar_full[stage]);
        }
        const int ti = static_cast<int>(kTapD[tap]), tj = static_cast<int>(kTapH[tap]), tk = static_cast<int>(kTapW[tap]);
        const float fdz = static_cast<float>(ud + ti);
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
          const int r = ps * kVoxPerPass + vsub;             // row inside the work unit (0..255)
          const float pd = fdz + ocur[ps][0];
          const float phh = static_cast<float>(vh[ps] + tj) + ocur[ps][1];
          const float pw = static_cast<float>(vw[ps] + tk) + ocur[ps][2];
          const bool inside = vlive[ps] && pd > -1.f && phh > -1.f && pw > -1.f && pd < static_cast<float>(D) &&
                              phh < static_cast<float>(H) && pw < static_cast<float>(W);
          // warp-uniform skip: when none of the warp's 8 voxels samples inside the volume for this tap, the A rows are zero and
          // neither the address arithmetic nor the loads are issued
          __nv_bfloat162 acc[8];
          if (!p.skip_oob || __any_sync(0xffffffffu, inside)) {
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
            const uint32_t b00 = (p.skip_oob && !inside) ? lane_off : static_cast<uint32_t>(vbase + dc0 * HW + hc0 * W) * cs2 + lane_off;
            const uint32_t b01 = (p.skip_oob && !inside) ? lane_off : static_cast<uint32_t>(vbase + dc0 * HW + hc1 * W) * cs2 + lane_off;
            const uint32_t b10 = (p.skip_oob && !inside) ? lane_off : static_cast<uint32_t>(vbase + dc1 * HW + hc0 * W) * cs2 + lane_off;
            const uint32_t b11 = (p.skip_oob && !inside) ? lane_off : static_cast<uint32_t>(vbase + dc1 * HW + hc1 * W) * cs2 + lane_off;
            // a sample outside the volume contributes zero and is not read (deform_im2col_cuda.cuh:248): its lanes point all eight
            // loads at ONE dummy line (voxel 0), so they add a single wavefront to the instruction instead of eight
            const uint32_t o0 = (inside || !p.skip_oob) ? static_cast<uint32_t>(wc0) * cs2 : 0u, o1 = (inside || !p.skip_oob) ? static_cast<uint32_t>(wc1) * cs2 : 0u;
            uint4 ua[8], ub2[8];
            ld_global_v8(ptr_add_u32(xbytes, b00 + o0), ua[0], ub2[0]);
            ld_global_v8(ptr_add_u32(xbytes, b00 + o1), ua[1], ub2[1]);
            ld_global_v8(ptr_add_u32(xbytes, b01 + o0), ua[2], ub2[2]);
            ld_global_v8(ptr_add_u32(xbytes, b01 + o1), ua[3], ub2[3]);
            ld_global_v8(ptr_add_u32(xbytes, b10 + o0), ua[4], ub2[4]);
            ld_global_v8(ptr_add_u32(xbytes, b10 + o1), ua[5], ub2[5]);
            ld_global_v8(ptr_add_u32(xbytes, b11 + o0), ua[6], ub2[6]);
            ld_global_v8(ptr_add_u32(xbytes, b11 + o1), ua[7], ub2[7]);
            const float a00 = wd0 * wh0, a01 = wd0 * wh1, a10 = wd1 * wh0, a11 = wd1 * wh1;
            const float cw[8] = {a00 * ww0, a00 * ww1, a01 * ww0, a01 * ww1, a10 * ww0, a10 * ww1, a11 * ww0, a11 * ww1};
            // packed bf16 blend (HFMA2.BF16): the blended A tile is rounded to bf16 for the MMA anyway
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
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = __float2bfloat162_rn(0.f);
          }
          if (lane_live) {
            uint4 oa, ob;
            oa.x = *reinterpret_cast<uint32_t*>(&acc[0]); oa.y = *reinterpret_cast<uint32_t*>(&acc[1]);
            oa.z = *reinterpret_cast<uint32_t*>(&acc[2]); oa.w = *reinterpret_cast<uint32_t*>(&acc[3]);
            ob.x = *reinterpret_cast<uint32_t*>(&acc[4]); ob.y = *reinterpret_cast<uint32_t*>(&acc[5]);
            ob.z = *reinterpret_cast<uint32_t*>(&acc[6]); ob.w = *reinterpret_cast<uint32_t*>(&acc[7]);
            const int blk = r >> 7, row = r & 127;
            uint8_t* dst = sa + blk * C::A_BLOCK_BYTES + (2 * cq) * C::A_CHUNK_BYTES + row * 16;
            *reinterpret_cast<uint4*>(dst) = oa;
            *reinterpret_cast<uint4*>(dst + C::A_CHUNK_BYTES) = ob;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[stage]);
      }
      if (STAGED && unit + 1 < unit_hi) {
        // half A of the next unit has landed; every producer is past its last read of half B: refill B for the next unit
        cp_async_wait<0>();
        prod_barrier();
        stage_offsets<kOffPiecesB, 3 * kSplitTap, kOffPitchB>(p, unit + 1, ptid, s_offB);
      }
      gbase += kTaps;
    }
  } else if (warp == kMmaWarp) {
    // ======================================= MMA issuer =====================================================
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, kNOut);
    const uint64_t adesc_hi = umma_desc_nosw(0, C::A_CHUNK_BYTES, 128);   // LBO = chunk pitch, SBO = 8 rows
    const uint64_t bdesc_hi = umma_desc_nosw(0, kNOut * 16, 128);
    const uint32_t sbase = smem_u32(s_stage);
    const bool leader = elect_one();
    uint32_t g = 0, it = 0;
    for (int unit = unit_lo; unit < unit_hi; ++unit, ++it) {
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
    for (int unit = unit_lo; unit < unit_hi; ++unit, ++it) {
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      mbar_wait(&bar_tfull[as], aph);      // (a nanosleep back-off here was measured 2 % slower: it delays the accumulator hand-back)
      tc_fence_after_sync();
      int ud, uh0, uw0, ub;
      unit_coords(unit, p, ud, uh0, uw0, ub);
#pragma unroll 1
      for (int blk = 0; blk < kBlocks; ++blk) {
        const int r = blk * 128 + warp * 32 + lane;
        const int hh = uh0 + (r >> 4), ww = uw0 + (r & 15);
        const long long v = (hh < H && ww < W) ? ((static_cast<long long>(ub) * D + ud) * H + hh) * W + ww : p.nvox;
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

template <int CINP, int OMODE>
int launch_dcn(const DcnParams& p, cudaStream_t st) {
  using C = DCfg<CINP, OMODE>;
  auto kern = dcn3d_kernel<CINP, OMODE>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return dpf::fail("dpf_dcn3d_fwd: cannot opt in to %d B shared memory: %s", C::SMEM_BYTES, cudaGetErrorString(e));
    // the gather lives on the L1 cache: ask for the smallest shared-memory carve-out that holds this kernel's buffers
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (C::SMEM_BYTES + 1024) * 100 / (228 * 1024) + 1);
    attr_done = true;
  }
  const int grid = std::min(p.nunits, dpf::sm_count());
  kern<<<grid, kThreads, C::SMEM_BYTES, st>>>(p);
  return dpf::after_launch("dpf_dcn3d_fwd");
}

}  // namespace

extern "C" int dpf_dcn3d_fwd(const void* x, const float* offset, const void* w, const float* scale, const float* shift,
                             void* y, int B, int D, int H, int W, int Cin_pad, int x_cstride, int off_cstride, int Cout, int relu,
                             void* stream) {
  DPF_REQUIRE(x && offset && w && y, "dpf_dcn3d_fwd: null pointer");
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(w) && DPF_ALIGNED16(y), "dpf_dcn3d_fwd: pointers must be 16-byte aligned");
  DPF_REQUIRE(Cout == kNOut, "dpf_dcn3d_fwd: Cout=%d, only 64 is built", Cout);
  DPF_REQUIRE(Cin_pad == 32 || Cin_pad == 64, "dpf_dcn3d_fwd: Cin_pad=%d must be 32 or 64", Cin_pad);
  DPF_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "dpf_dcn3d_fwd: bad shape");
  DPF_REQUIRE(static_cast<long long>(B) * D * H * W < (1LL << 31) / 128, "dpf_dcn3d_fwd: tensor too large for 32-bit voxel indexing");
  DPF_REQUIRE(x_cstride >= Cin_pad && x_cstride % 8 == 0, "dpf_dcn3d_fwd: x_cstride=%d must be a multiple of 8 >= Cin_pad", x_cstride);
  DPF_REQUIRE(off_cstride >= 81, "dpf_dcn3d_fwd: off_cstride=%d must be >= 81", off_cstride);
  DcnParams p{};
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.offset = offset;
  p.off_cstride = off_cstride;
  p.w = reinterpret_cast<const __nv_bfloat16*>(w);
  p.scale = scale;
  p.shift = shift;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.B = B; p.D = D; p.H = H; p.W = W; p.relu = relu; p.x_cstride = x_cstride;
  p.nvox = static_cast<long long>(B) * D * H * W;
  p.tiles_h = (H + kUnitH - 1) / kUnitH;
  p.tiles_w = (W + 15) / 16;
  p.nunits = B * D * p.tiles_h * p.tiles_w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // offset path: 2 (default) = 16-byte loads + shuffles, 1 = staged in shared memory, 0 = scalar loads.  Modes 1 and 2 need
  // 16-byte aligned voxel rows that hold floats 0..83 (the 81 real offsets + padding); anything else takes mode 0.
  {
    static int skip_env = -1;
    if (skip_env < 0) { const char* e = getenv("DPF_DCN_SKIP"); skip_env = e ? atoi(e) : 1; }
    p.skip_oob = skip_env;
  }
  static int omode_env = -1;
  if (omode_env < 0) { const char* e = getenv("DPF_DCN_OMODE"); omode_env = e ? atoi(e) : 2; }
  const bool vec_ok = off_cstride % 4 == 0 && off_cstride >= 84 && (reinterpret_cast<uintptr_t>(offset) & 15u) == 0;
  const int omode = vec_ok ? omode_env : 0;
  if (Cin_pad == 32) return omode == 2 ? launch_dcn<32, 2>(p, st) : omode == 1 ? launch_dcn<32, 1>(p, st) : launch_dcn<32, 0>(p, st);
  return omode == 2 ? launch_dcn<64, 2>(p, st) : omode == 1 ? launch_dcn<64, 1>(p, st) : launch_dcn<64, 0>(p, st);
}
