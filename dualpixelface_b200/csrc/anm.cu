// ANM front end (normal branch) for sm_100a: level selection, coordinate volume, feature-volume assembly.
//
// Replaces, from src/model/stereodpnet/normal_module.py of the reference:
//   :156      F.interpolate(disp, 0.25, 'nearest') * 0.25        -> pixel (4h, 4w) of the full-res disparity
//   :130-136  sample_with_sort: topk(1/(|level-d|+1e-6), K) -> sort -> gather
//   :80-118   grid_maker_3d: K^-1 [u,v,1] * depth(disp) with depth = a/(d-b) (src/utils/geometry.py:35-40),
//             per-sample min/max normalisation with +1e-6
//   :166      cat([cost, coord]) -> [B, C+3, K, H4, W4]  (here [B,K,H4,W4,Cpad] bf16, channels-last, zero padded)
// Index selection is integer-exact.  Tie rule (d exactly on a level makes the K-th candidate a two-way tie, where
// torch.topk's winner is implementation-defined): the LOWER level index wins.
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"
#include <cfloat>

namespace {

constexpr int kMaxLevels = 16;
constexpr int kMaxK = 8;

struct Levels {
  float v[kMaxLevels];
};

__device__ __forceinline__ void atomic_min_f(float* addr, float val) {
  if (val >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(val));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(val));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float val) {
  if (val >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(val));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(val));
}

__global__ void __launch_bounds__(256) anm_select_kernel(const float* __restrict__ disp, const float* __restrict__ kinv,
                                                         const float* __restrict__ abvalue, Levels lv, int* __restrict__ idx,
                                                         float* __restrict__ coord, float* __restrict__ minmax, int D, int K,
                                                         int H4, int W4) {
  const int b = blockIdx.y;
  const int n = H4 * W4;
  const int H = 4 * H4, W = 4 * W4;
  float lmin = FLT_MAX, lmax = -FLT_MAX;
  const float a = abvalue[b * 2 + 1], bb = abvalue[b * 2 + 0];
  float ki[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) ki[i] = kinv[b * 9 + i];
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < n; pix += gridDim.x * blockDim.x) {
    const int h = pix / W4, w = pix - h * W4;
    const float d = disp[(static_cast<size_t>(b) * H + 4 * h) * W + 4 * w] * 0.25f;
    float score[kMaxLevels];
#pragma unroll
    for (int l = 0; l < kMaxLevels; ++l) score[l] = (l < D) ? 1.0f / (fabsf(lv.v[l] - d) + 1e-6f) : -1.0f;
    unsigned taken = 0u;
    for (int k = 0; k < K; ++k) {                       // top-K by repeated arg-max, strict '>' => lower index wins ties
      int best = 0;
      float bs = -2.0f;
#pragma unroll
      for (int l = 0; l < kMaxLevels; ++l) {
        const bool free_l = ((taken >> l) & 1u) == 0u;
        if (free_l && score[l] > bs) { bs = score[l]; best = l; }
      }
      taken |= 1u << best;
    }
    const float u = static_cast<float>(w), v = static_cast<float>(h);
    const float rx = ki[0] * u + ki[1] * v + ki[2];
    const float ry = ki[3] * u + ki[4] * v + ki[5];
    const float rz = ki[6] * u + ki[7] * v + ki[8];
    int k = 0;
    for (int l = 0; l < D; ++l) {                       // ascending level order == torch.sort of the indices
      if (((taken >> l) & 1u) == 0u) continue;
      idx[(static_cast<size_t>(b) * K + k) * n + pix] = l;
      float depth = a / (lv.v[l] - bb);
      if (isnan(depth) || isinf(depth)) depth = 0.f;
      const float cx = rx * depth, cy = ry * depth, cz = rz * depth;
      float* co = coord + ((static_cast<size_t>(b) * K + k) * n + pix) * 3;
      co[0] = cx; co[1] = cy; co[2] = cz;
      lmin = fminf(lmin, fminf(cx, fminf(cy, cz)));
      lmax = fmaxf(lmax, fmaxf(cx, fmaxf(cy, cz)));
      ++k;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if ((threadIdx.x & 31) == 0 && lmin <= lmax) {
    atomic_min_f(&minmax[b * 2 + 0], lmin);
    atomic_max_f(&minmax[b * 2 + 1], lmax);
  }
}

__global__ void __launch_bounds__(256) anm_gather_kernel(const __nv_bfloat16* __restrict__ out3, const int* __restrict__ idx,
                                                         const float* __restrict__ coord, const float* __restrict__ minmax,
                                                         __nv_bfloat16* __restrict__ fv, int B, int D, int K, int H4, int W4,
                                                         int C, int Cpad) {
  const int pcs = Cpad >> 3;
  const size_t n = static_cast<size_t>(H4) * W4;
  const long long total = static_cast<long long>(B) * K * n * pcs;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pc = static_cast<int>(q % pcs);
    long long t = q / pcs;
    const size_t pix = static_cast<size_t>(t % n);
    t /= n;
    const int k = static_cast<int>(t % K);
    const int b = static_cast<int>(t / K);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    const int c0 = pc * 8;
    if (c0 + 8 <= C) {
      const int l = idx[(static_cast<size_t>(b) * K + k) * n + pix];
      o = __ldg(reinterpret_cast<const uint4*>(out3 + ((static_cast<size_t>(b) * D + l) * n + pix) * C + c0));
    } else if (c0 == C) {
      const float mn = minmax[b * 2 + 0], mx = minmax[b * 2 + 1];
      const float inv = 1.0f / (mx - mn + 1e-6f);
      const float* co = coord + ((static_cast<size_t>(b) * K + k) * n + pix) * 3;
      o.x = dpf::pack_bf16x2((co[0] - mn) * inv, (co[1] - mn) * inv);
      o.y = dpf::pack_bf16x2((co[2] - mn) * inv, 0.f);
    }
    *reinterpret_cast<uint4*>(fv + ((static_cast<size_t>(b) * K + k) * n + pix) * Cpad + c0) = o;
  }
}

}  // namespace

extern "C" int dpf_anm_select(const float* disp, const float* kinv, const float* abvalue, const float* levels_host, int* idx,
                              float* coord, float* minmax, int B, int D, int K, int H4, int W4, void* stream) {
  DPF_REQUIRE(disp && kinv && abvalue && levels_host && idx && coord && minmax, "dpf_anm_select: null pointer");
  DPF_REQUIRE(D >= 1 && D <= kMaxLevels && K >= 1 && K <= kMaxK && K <= D, "dpf_anm_select: bad D=%d K=%d", D, K);
  DPF_REQUIRE(B > 0 && B <= 65535 && H4 > 0 && W4 > 0, "dpf_anm_select: bad shape");
  Levels lv;
  for (int i = 0; i < kMaxLevels; ++i) lv.v[i] = i < D ? levels_host[i] : 0.f;
  const int n = H4 * W4;
  const int blocks = std::min((n + 255) / 256, dpf::sm_count() * 8 / B + 1);
  anm_select_kernel<<<dim3(blocks, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(disp, kinv, abvalue, lv, idx, coord,
                                                                                   minmax, D, K, H4, W4);
  return dpf::after_launch("dpf_anm_select");
}

extern "C" int dpf_anm_gather(const void* out3, const int* idx, const float* coord, const float* minmax, void* fv, int B,
                              int D, int K, int H4, int W4, int C, int Cpad, void* stream) {
  DPF_REQUIRE(out3 && idx && coord && minmax && fv, "dpf_anm_gather: null pointer");
  DPF_REQUIRE(DPF_ALIGNED16(out3) && DPF_ALIGNED16(fv), "dpf_anm_gather: pointers must be 16-byte aligned");
  DPF_REQUIRE(C % 8 == 0 && Cpad % 8 == 0 && Cpad >= C + 8, "dpf_anm_gather: need C %% 8 == 0 and Cpad >= C + 8 (C=%d Cpad=%d)", C, Cpad);
  const long long total = static_cast<long long>(B) * K * H4 * W4 * (Cpad / 8);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(dpf::sm_count()) * 16));
  anm_gather_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(out3), idx, coord, minmax, reinterpret_cast<__nv_bfloat16*>(fv), B, D, K, H4, W4, C, Cpad);
  return dpf::after_launch("dpf_anm_gather");
}
