// Integer-shift cost-volume builder (memory-bound) for sm_100a.
//
// Replaces CostVolume.build_concat_volume / build_gwc_volume of the reference (src/model/psmnet/modules.py:223-262)
// and the difference volume of src/model/stereonet/mainmodel.py:100-114:
//   vol[b,i,h,w,:] = f(ref[b,h,w,:], tgt[b,h+s_i,w,:])  if 0 <= h+s_i < H  else 0
// Layout: features [B,H,W,C] bf16 (NHWC), volume [B,D,H,W,Cv] bf16 (NDHWC).
//
// One CTA owns an R-row x WS-pixel strip of one image.  The ref rows and the tgt rows h0+smin .. h0+R-1+smax are
// staged ONCE in shared memory with 1-D TMA bulk copies (cp.async.bulk, mbarrier complete_tx), then every
// disparity level is emitted from shared memory with fully coalesced 128-bit streaming stores (a warp writes
// 512 contiguous bytes per instruction).  Algorithmic traffic: read 2*C*2 B, write D*Cv*2 B per pixel.
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"

namespace {

constexpr int kRows = 4;       // output rows per CTA (45 KB of staging -> 5 CTAs per SM overlap their load and store phases)
constexpr int kWS = 64;        // pixels per strip
constexpr int kThreads = 256;
constexpr int kMaxD = 16;

struct Shifts {
  int s[kMaxD];
  int smin, smax;
};

__device__ __forceinline__ uint4 zero4() { return make_uint4(0u, 0u, 0u, 0u); }

__device__ __forceinline__ uint4 sub_bf16x8(const uint4& a, const uint4& b) {
  uint4 r;
  r.x = dpf::pack_bf16x2(dpf::bf16_lo(a.x) - dpf::bf16_lo(b.x), dpf::bf16_hi(a.x) - dpf::bf16_hi(b.x));
  r.y = dpf::pack_bf16x2(dpf::bf16_lo(a.y) - dpf::bf16_lo(b.y), dpf::bf16_hi(a.y) - dpf::bf16_hi(b.y));
  r.z = dpf::pack_bf16x2(dpf::bf16_lo(a.z) - dpf::bf16_lo(b.z), dpf::bf16_hi(a.z) - dpf::bf16_hi(b.z));
  r.w = dpf::pack_bf16x2(dpf::bf16_lo(a.w) - dpf::bf16_lo(b.w), dpf::bf16_hi(a.w) - dpf::bf16_hi(b.w));
  return r;
}

// MODE 0 concat, 1 difference, 2 group-wise correlation
template <int MODE>
__global__ void __launch_bounds__(kThreads) costvol_fwd_kernel(const __nv_bfloat16* __restrict__ ref,
                                                               const __nv_bfloat16* __restrict__ tgt,
                                                               __nv_bfloat16* __restrict__ vol, int H, int W, int C, int D,
                                                               int G, Shifts sh) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  const int w0 = blockIdx.x * kWS;
  const int h0 = blockIdx.y * kRows;
  const int b = blockIdx.z;
  const int wseg = min(kWS, W - w0);
  const int rows = min(kRows, H - h0);
  const int rowBytes = kWS * C * 2;                 // smem pitch of one staged row
  const int segBytes = wseg * C * 2;
  const int tRows = kRows + sh.smax - sh.smin;
  uint8_t* sref = smem;
  uint8_t* stgt = smem + kRows * rowBytes;

  if (threadIdx.x == 0) {
    dpf::mbar_init(&bar, 1);
    dpf::mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nvalid = rows;
    for (int j = 0; j < tRows; ++j) {
      const int h = h0 + sh.smin + j;
      if (h >= 0 && h < H && j < rows + sh.smax - sh.smin) ++nvalid;
    }
    dpf::mbar_arrive_expect_tx(&bar, static_cast<uint32_t>(nvalid) * segBytes);
    const size_t img = static_cast<size_t>(b) * H;
    for (int r = 0; r < rows; ++r)
      dpf::bulk_g2s(dpf::smem_u32(sref + r * rowBytes), ref + ((img + h0 + r) * W + w0) * C, segBytes, &bar);
    for (int j = 0; j < tRows; ++j) {
      const int h = h0 + sh.smin + j;
      if (h >= 0 && h < H && j < rows + sh.smax - sh.smin)
        dpf::bulk_g2s(dpf::smem_u32(stgt + j * rowBytes), tgt + ((img + h) * W + w0) * C, segBytes, &bar);
    }
  }
  dpf::mbar_wait(&bar, 0);

  const int c8 = C >> 3;                                        // 16-byte pieces per feature pixel
  const int pp = (MODE == 0) ? 2 * c8 : (MODE == 1 ? c8 : (G >> 3));   // 16-byte pieces per volume pixel
  const int cv = pp * 8;
  const int piecesPerRow = wseg * pp;
  const int total = rows * piecesPerRow;
  const float inv_cpg = (MODE == 2) ? -1.0f / static_cast<float>(C / G) : 0.f;

  if (MODE == 2) {
    // Group-wise correlation (build_gwc_volume, psmnet/modules.py:215-221,243-262): vol[g] = -mean_{c in group g}(ref[c]*tgt[c]).
    // C/8 lanes per pixel: a lane loads ONE 16-byte piece of the ref pixel and of the shifted tgt pixel from the staged rows
    // (conflict-free 128-bit shared loads), multiplies its 8 channels and sums them per group in registers; groups wider than
    // 8 channels are finished by a warp-shuffle (xor) reduction over the group's lanes; the G results of a pixel are then
    // collected by shuffles into the lanes that own the pixel's 16-byte output pieces and stored with coalesced 128-bit
    // streaming stores.  (Round 1 walked the channels with scalar 2-byte shared loads in one thread per output piece: 266 GB/s.)
    const int lpp = c8;                                          // lanes per pixel (power of two <= 32: C in {8,16,...,256})
    const int cpg = C / G;                                       // channels per group
    const int ng = (cpg < 8) ? 8 / cpg : 1;                      // group sums a lane holds
    const int lpg = (cpg > 8) ? cpg / 8 : 1;                     // lanes per group
    const int totalLanes = rows * wseg * lpp;
    const int totalPad = (totalLanes + 31) & ~31;
    for (int i = 0; i < D; ++i) {
      const int s = sh.s[i];
      __nv_bfloat16* vbase = vol + ((static_cast<size_t>(b) * D + i) * H + h0) * static_cast<size_t>(W) * cv;
      for (int q = threadIdx.x; q < totalPad; q += kThreads) {
        const bool live = q < totalLanes;
        const int qq = live ? q : 0;
        const int pc = qq % lpp;                                 // this lane's 16-byte input piece == the output piece it may own
        const int pix = qq / lpp;
        const int r = pix / wseg, px = pix - r * wseg;
        const int hs = h0 + r + s;
        const bool valid = live && (hs >= 0) && (hs < H);
        float gs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (valid) {
          const uint4 a = *reinterpret_cast<const uint4*>(sref + r * rowBytes + px * (C * 2) + pc * 16);
          const uint4 t = *reinterpret_cast<const uint4*>(stgt + (r + s - sh.smin) * rowBytes + px * (C * 2) + pc * 16);
          const float pr[8] = {dpf::bf16_lo(a.x) * dpf::bf16_lo(t.x), dpf::bf16_hi(a.x) * dpf::bf16_hi(t.x),
                               dpf::bf16_lo(a.y) * dpf::bf16_lo(t.y), dpf::bf16_hi(a.y) * dpf::bf16_hi(t.y),
                               dpf::bf16_lo(a.z) * dpf::bf16_lo(t.z), dpf::bf16_hi(a.z) * dpf::bf16_hi(t.z),
                               dpf::bf16_lo(a.w) * dpf::bf16_lo(t.w), dpf::bf16_hi(a.w) * dpf::bf16_hi(t.w)};
          if (cpg >= 8) {
            gs[0] = ((pr[0] + pr[1]) + (pr[2] + pr[3])) + ((pr[4] + pr[5]) + (pr[6] + pr[7]));
          } else if (cpg == 4) {
            gs[0] = (pr[0] + pr[1]) + (pr[2] + pr[3]); gs[1] = (pr[4] + pr[5]) + (pr[6] + pr[7]);
          } else if (cpg == 2) {
            gs[0] = pr[0] + pr[1]; gs[1] = pr[2] + pr[3]; gs[2] = pr[4] + pr[5]; gs[3] = pr[6] + pr[7];
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) gs[k] = pr[k];
          }
        }
        for (int o = 1; o < lpg; o <<= 1) gs[0] += __shfl_xor_sync(0xffffffffu, gs[0], o);     // groups wider than one lane
        // lane pc of the pixel assembles output piece pc = groups 8*pc .. 8*pc+7; group g lives in lane g*lpg (slot 0) when
        // cpg >= 8, else in lane g / ng, slot g % ng = j % ng (ng divides 8): the slot is the same for every reader
        const int lane = threadIdx.x & 31;
        const int pixbase = lane - pc;
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int g = 8 * pc + j;
          const int src = pixbase + ((cpg >= 8) ? g * lpg : g / ng);
          const float mine = (ng == 1) ? gs[0] : (ng == 2 ? gs[j & 1] : (ng == 4 ? gs[j & 3] : gs[j]));
          o8[j] = __shfl_sync(0xffffffffu, mine, src & 31) * inv_cpg;
        }
        if (live && pc < pp) {
          uint4 v;
          v.x = dpf::pack_bf16x2(o8[0], o8[1]); v.y = dpf::pack_bf16x2(o8[2], o8[3]);
          v.z = dpf::pack_bf16x2(o8[4], o8[5]); v.w = dpf::pack_bf16x2(o8[6], o8[7]);
          dpf::st_cs_v4(vbase + (static_cast<size_t>(r) * W + w0 + px) * cv + pc * 8, v);
        }
      }
    }
    return;
  }

  for (int i = 0; i < D; ++i) {
    const int s = sh.s[i];
    __nv_bfloat16* vbase = vol + ((static_cast<size_t>(b) * D + i) * H + h0) * static_cast<size_t>(W) * cv;
#pragma unroll 4
    for (int q = threadIdx.x; q < total; q += kThreads) {
      const int r = q / piecesPerRow;
      const int rem = q - r * piecesPerRow;
      const int px = rem / pp;
      const int pc = rem - px * pp;
      const int hs = h0 + r + s;
      const bool valid = (hs >= 0) && (hs < H);
      uint4 v = zero4();
      if (valid) {
        const uint8_t* rrow = sref + r * rowBytes + px * (C * 2);
        const uint8_t* trow = stgt + (r + s - sh.smin) * rowBytes + px * (C * 2);
        if (MODE == 0) {
          v = (pc < c8) ? *reinterpret_cast<const uint4*>(rrow + pc * 16)
                        : *reinterpret_cast<const uint4*>(trow + (pc - c8) * 16);
        } else {
          v = sub_bf16x8(*reinterpret_cast<const uint4*>(rrow + pc * 16), *reinterpret_cast<const uint4*>(trow + pc * 16));
        }
      }
      dpf::st_cs_v4(vbase + (static_cast<size_t>(r) * W + w0 + px) * cv + pc * 8, v);
    }
  }
}

// Gradient wrt ref / tgt in gather form: one thread per 16-byte piece of dref or dtgt.
template <int MODE>
__global__ void __launch_bounds__(256) costvol_bwd_kernel(const __nv_bfloat16* __restrict__ ref,
                                                          const __nv_bfloat16* __restrict__ tgt,
                                                          const __nv_bfloat16* __restrict__ dvol,
                                                          __nv_bfloat16* __restrict__ dref, __nv_bfloat16* __restrict__ dtgt,
                                                          int B, int H, int W, int C, int D, int G, Shifts sh) {
  const int c8 = C >> 3;
  const long long npieces = static_cast<long long>(B) * H * W * c8 * 2;
  const int cv = (MODE == 0) ? 2 * C : (MODE == 1 ? C : G);
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < npieces;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pc = static_cast<int>(q % c8);
    long long t = q / c8;
    const int side = static_cast<int>(t & 1);      // 0 -> dref, 1 -> dtgt
    t >>= 1;
    const int w = static_cast<int>(t % W);
    t /= W;
    const int h = static_cast<int>(t % H);
    const int b = static_cast<int>(t / H);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < D; ++i) {
      const int s = sh.s[i];
      const int hv = side ? (h - s) : h;             // volume row that holds this element's contribution
      const int hp = side ? hv : (h + s);            // partner row (tgt row for dref, ref row for dtgt)
      if (hv < 0 || hv >= H || hp < 0 || hp >= H) continue;
      const __nv_bfloat16* dv = dvol + (((static_cast<size_t>(b) * D + i) * H + hv) * W + w) * cv;
      if (MODE == 0 || MODE == 1) {
        const int coff = (MODE == 0 && side) ? C : 0;
        const uint4 g = *reinterpret_cast<const uint4*>(dv + coff + pc * 8);
        const float sgn = (MODE == 1 && side) ? -1.f : 1.f;
        acc[0] += sgn * dpf::bf16_lo(g.x); acc[1] += sgn * dpf::bf16_hi(g.x);
        acc[2] += sgn * dpf::bf16_lo(g.y); acc[3] += sgn * dpf::bf16_hi(g.y);
        acc[4] += sgn * dpf::bf16_lo(g.z); acc[5] += sgn * dpf::bf16_hi(g.z);
        acc[6] += sgn * dpf::bf16_lo(g.w); acc[7] += sgn * dpf::bf16_hi(g.w);
      } else {
        const int cpg = C / G;
        const __nv_bfloat16* other = (side ? ref : tgt) + ((static_cast<size_t>(b) * H + hp) * W + w) * C + pc * 8;
        const float k = -1.0f / static_cast<float>(cpg);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int g = (pc * 8 + c) / cpg;
          acc[c] += k * __bfloat162float(dv[g]) * __bfloat162float(other[c]);
        }
      }
    }
    uint4 o;
    o.x = dpf::pack_bf16x2(acc[0], acc[1]);
    o.y = dpf::pack_bf16x2(acc[2], acc[3]);
    o.z = dpf::pack_bf16x2(acc[4], acc[5]);
    o.w = dpf::pack_bf16x2(acc[6], acc[7]);
    __nv_bfloat16* dst = (side ? dtgt : dref) + ((static_cast<size_t>(b) * H + h) * W + w) * C + pc * 8;
    *reinterpret_cast<uint4*>(dst) = o;
  }
}

int make_shifts(const int* shifts_host, int D, Shifts* out) {
  out->smin = 0;
  out->smax = 0;
  for (int i = 0; i < kMaxD; ++i) out->s[i] = 0;
  for (int i = 0; i < D; ++i) {
    out->s[i] = shifts_host[i];
    if (shifts_host[i] < out->smin) out->smin = shifts_host[i];
    if (shifts_host[i] > out->smax) out->smax = shifts_host[i];
  }
  return 0;
}

}  // namespace

extern "C" int dpf_costvol_fwd(int mode, const void* ref, const void* tgt, void* vol, int B, int H4, int W4, int C, int D,
                               int G, const int* shifts_host, void* stream) {
  DPF_REQUIRE(mode >= 0 && mode <= 2, "dpf_costvol_fwd: mode %d not in {0,1,2}", mode);
  DPF_REQUIRE(ref && tgt && vol && shifts_host, "dpf_costvol_fwd: null pointer");
  DPF_REQUIRE(DPF_ALIGNED16(ref) && DPF_ALIGNED16(tgt) && DPF_ALIGNED16(vol), "dpf_costvol_fwd: pointers must be 16-byte aligned");
  DPF_REQUIRE(B > 0 && H4 > 0 && W4 > 0 && B <= 65535, "dpf_costvol_fwd: bad shape B=%d H4=%d W4=%d", B, H4, W4);
  DPF_REQUIRE(C >= 8 && C % 8 == 0 && C <= 128, "dpf_costvol_fwd: C=%d must be a multiple of 8 in [8,128]", C);
  DPF_REQUIRE(D >= 1 && D <= kMaxD, "dpf_costvol_fwd: D=%d must be in [1,%d]", D, kMaxD);
  if (mode == 2) DPF_REQUIRE(G >= 8 && G % 8 == 0 && C % G == 0 && ((C / 8) & (C / 8 - 1)) == 0 && ((C / G) & (C / G - 1)) == 0,
                             "dpf_costvol_fwd: gwc needs G %% 8 == 0, G | C, and C/8, C/G powers of two (G=%d C=%d)", G, C);
  Shifts sh;
  make_shifts(shifts_host, D, &sh);
  DPF_REQUIRE(sh.smax - sh.smin <= 16, "dpf_costvol_fwd: shift span %d too large", sh.smax - sh.smin);
  const size_t smem = static_cast<size_t>(2 * kRows + sh.smax - sh.smin) * kWS * C * 2;
  dim3 grid((W4 + kWS - 1) / kWS, (H4 + kRows - 1) / kRows, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto r = reinterpret_cast<const __nv_bfloat16*>(ref);
  auto t = reinterpret_cast<const __nv_bfloat16*>(tgt);
  auto v = reinterpret_cast<__nv_bfloat16*>(vol);
  cudaError_t e = cudaSuccess;
  if (mode == 0) {
    e = cudaFuncSetAttribute(costvol_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) costvol_fwd_kernel<0><<<grid, kThreads, smem, st>>>(r, t, v, H4, W4, C, D, G, sh);
  } else if (mode == 1) {
    e = cudaFuncSetAttribute(costvol_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) costvol_fwd_kernel<1><<<grid, kThreads, smem, st>>>(r, t, v, H4, W4, C, D, G, sh);
  } else {
    e = cudaFuncSetAttribute(costvol_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) costvol_fwd_kernel<2><<<grid, kThreads, smem, st>>>(r, t, v, H4, W4, C, D, G, sh);
  }
  if (e != cudaSuccess) return dpf::fail("dpf_costvol_fwd: %s", cudaGetErrorString(e));
  return dpf::after_launch("dpf_costvol_fwd");
}

extern "C" int dpf_costvol_bwd(int mode, const void* ref, const void* tgt, const void* dvol, void* dref, void* dtgt, int B,
                               int H4, int W4, int C, int D, int G, const int* shifts_host, void* stream) {
  DPF_REQUIRE(mode >= 0 && mode <= 2, "dpf_costvol_bwd: mode %d not in {0,1,2}", mode);
  DPF_REQUIRE(dvol && dref && dtgt && shifts_host, "dpf_costvol_bwd: null pointer");
  DPF_REQUIRE(mode != 2 || (ref && tgt), "dpf_costvol_bwd: gwc needs ref and tgt");
  DPF_REQUIRE(DPF_ALIGNED16(dvol) && DPF_ALIGNED16(dref) && DPF_ALIGNED16(dtgt), "dpf_costvol_bwd: pointers must be 16-byte aligned");
  DPF_REQUIRE(C >= 8 && C % 8 == 0, "dpf_costvol_bwd: C=%d must be a multiple of 8", C);
  DPF_REQUIRE(D >= 1 && D <= kMaxD, "dpf_costvol_bwd: D=%d must be in [1,%d]", D, kMaxD);
  if (mode == 2) DPF_REQUIRE(G >= 8 && G % 8 == 0 && C % G == 0, "dpf_costvol_bwd: gwc needs G %% 8 == 0 and G | C");
  Shifts sh;
  make_shifts(shifts_host, D, &sh);
  const long long npieces = static_cast<long long>(B) * H4 * W4 * (C / 8) * 2;
  const int blocks = static_cast<int>(std::min<long long>((npieces + 255) / 256, static_cast<long long>(dpf::sm_count()) * 16));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto r = reinterpret_cast<const __nv_bfloat16*>(ref);
  auto t = reinterpret_cast<const __nv_bfloat16*>(tgt);
  auto dv = reinterpret_cast<const __nv_bfloat16*>(dvol);
  auto dr = reinterpret_cast<__nv_bfloat16*>(dref);
  auto dt = reinterpret_cast<__nv_bfloat16*>(dtgt);
  if (mode == 0) costvol_bwd_kernel<0><<<blocks, 256, 0, st>>>(r, t, dv, dr, dt, B, H4, W4, C, D, G, sh);
  else if (mode == 1) costvol_bwd_kernel<1><<<blocks, 256, 0, st>>>(r, t, dv, dr, dt, B, H4, W4, C, D, G, sh);
  else costvol_bwd_kernel<2><<<blocks, 256, 0, st>>>(r, t, dv, dr, dt, B, H4, W4, C, D, G, sh);
  return dpf::after_launch("dpf_costvol_bwd");
}
