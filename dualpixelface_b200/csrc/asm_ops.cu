// ASM (adaptive sampling module) memory-bound pieces of the StereoDPNet cost volume, sm_100a.
//
//  dpf_asm_sample_fwd : table-driven separable resampling = subpixel_shift.forward of the reference
//                       (src/module/asm/asm.py:87-127).  The host builds per-row / per-column (index, weight) tables
//                       with the reference's own fp32 op sequence (grid normalise -> grid_sample un-normalise), so the
//                       sampling coordinates are bit-exact by construction; the kernel only gathers and blends.
//  dpf_channel_stats  : per-(b,c) sum / sum of squares (InstanceNorm3d statistics of the mask logits, asm.py:138).
//  dpf_asm_blend_fwd  : sigmoid -> softmax over the S samples -> weighted mean (asm.py:160-171) written straight into
//                       the [B,D,H4,W4,2C] bf16 volume slices (src/model/stereodpnet/modules.py:193-194).
#include "../../include/dpf_sm100.h"
#include "dpf_common.cuh"
#include "dpf_ptx.cuh"

namespace {

using namespace dpf;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

// out [B,S,H,W,C]
__global__ void __launch_bounds__(256) asm_sample_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                         int B, int H, int W, int C, int S, const int* __restrict__ ri,
                                                         const float* __restrict__ rw, const int* __restrict__ ci,
                                                         const float* __restrict__ cw) {
  const int c8n = C >> 3;
  const long long total = static_cast<long long>(B) * S * H * W * c8n;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(q % c8n);
    long long t = q / c8n;
    const int w = static_cast<int>(t % W); t /= W;
    const int h = static_cast<int>(t % H); t /= H;
    const int s = static_cast<int>(t % S);
    const int b = static_cast<int>(t / S);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = ri[(s * H + h) * 2 + i];
      const float wr = rw[(s * H + h) * 2 + i];
      if (r < 0 || wr == 0.f) continue;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = ci[(s * W + w) * 2 + j];
        const float wc = cw[(s * W + w) * 2 + j];
        if (c < 0 || wc == 0.f) continue;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(b) * H + r) * W + c) * C) + c8);
        float f[8];
        unpack8(u, f);
        const float ww = wr * wc;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(ww, f[k], acc[k]);
      }
    }
    st_cs_v4(out + (((static_cast<size_t>(b) * S + s) * H + h) * W + w) * C + c8 * 8, pack8(acc));
  }
}

// x [B,P,C] bf16 -> stats [B,C,2] (sum, sumsq), DETERMINISTIC: no floating-point atomics anywhere.
//   pass 1, grid (chunks, B), block 256 = (256/c8n) position lanes x c8n 16-byte pieces: every thread sums its positions in a
//           fixed order, the block combines its lanes in lane order through shared memory and writes ONE partial row
//           ws[b][chunk][2C];
//   pass 2, grid B, block 2C: thread i adds the chunk partials of its column in chunk order.
// The summation tree depends only on (P, C, chunks), never on scheduling, so the InstanceNorm / BatchNorm statistics -- and
// everything downstream of them -- are bit-identical from run to run.
__global__ void __launch_bounds__(256) channel_stats_partial_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ ws,
                                                                    long long P, int C) {
  extern __shared__ float part[];                // [lanes][2C]
  const int c8n = C >> 3;
  const int b = blockIdx.y;
  const int piece = threadIdx.x % c8n;
  const int lanes = blockDim.x / c8n;
  const int pl = threadIdx.x / c8n;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ss[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (pl < lanes) {
    const __nv_bfloat16* xb = x + static_cast<size_t>(b) * P * C;
    for (long long pos = static_cast<long long>(blockIdx.x) * lanes + pl; pos < P; pos += static_cast<long long>(gridDim.x) * lanes) {
      const uint4 u = ld_nc_v4(reinterpret_cast<const uint4*>(xb + pos * C) + piece);
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] += f[k]; ss[k] = fmaf(f[k], f[k], ss[k]); }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      part[pl * 2 * C + piece * 8 + k] = s[k];
      part[pl * 2 * C + C + piece * 8 + k] = ss[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += part[l * 2 * C + i];
    ws[(static_cast<size_t>(b) * gridDim.x + blockIdx.x) * 2 * C + i] = acc;
  }
}

__global__ void channel_stats_final_kernel(const float* __restrict__ ws, float* __restrict__ stats, int chunks, int C) {
  const int b = blockIdx.x, i = threadIdx.x;     // i in [0, 2C): column of the partial rows
  float acc = 0.f;
  for (int ch = 0; ch < chunks; ++ch) acc += ws[(static_cast<size_t>(b) * chunks + ch) * 2 * C + i];
  const int c = i < C ? i : i - C;
  stats[(static_cast<size_t>(b) * C + c) * 2 + (i < C ? 0 : 1)] = acc;
}

inline int channel_stats_chunks(int B, long long P, int C) {
  const int lanes = 256 / (C / 8);
  return static_cast<int>(std::min<long long>((P + lanes * 8 - 1) / (lanes * 8), static_cast<long long>(dpf::sm_count()) * 8 / B + 1));
}

// samples/logits [B,S,H,W,C] bf16; vol [B,D,H,W,Cvol]
template <int S>
__global__ void __launch_bounds__(256) asm_blend_kernel(const __nv_bfloat16* __restrict__ samples,
                                                        const __nv_bfloat16* __restrict__ logits,
                                                        const float* __restrict__ in_a, const float* __restrict__ in_d,
                                                        __nv_bfloat16* __restrict__ vol, int B, int H, int W, int C, int Dvol,
                                                        int d0, int D_rep, int ch_off, int Cvol) {
  const int c8n = C >> 3;
  const long long total = static_cast<long long>(B) * H * W * c8n;
  const size_t plane = static_cast<size_t>(H) * W;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(q % c8n);
    const long long pix = q / c8n;                       // b*H*W + h*W + w
    const int b = static_cast<int>(pix / plane);
    const size_t hw = static_cast<size_t>(pix % plane);
    float xs[S][8], gs[S][8];
    float a[8], dsh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a[k] = in_a[b * C + c8 * 8 + k];
      dsh[k] = in_d[b * C + c8 * 8 + k];
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const size_t off = ((static_cast<size_t>(b) * S + s) * plane + hw) * C + c8 * 8;
      unpack8(ld_nc_v4(samples + off), xs[s]);
      float l[8];
      unpack8(ld_nc_v4(logits + off), l);
#pragma unroll
      for (int k = 0; k < 8; ++k) gs[s][k] = 1.0f / (1.0f + __expf(-(l[k] * a[k] + dsh[k])));   // sigmoid(IN(logit))
    }
    float y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float m = gs[0][k];
#pragma unroll
      for (int s = 1; s < S; ++s) m = fmaxf(m, gs[s][k]);
      float den = 0.f, num = 0.f;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float e = __expf(gs[s][k] - m);
        den += e;
        num = fmaf(e, xs[s][k], num);
      }
      y[k] = num / (den * static_cast<float>(S));
    }
    const uint4 o = pack8(y);
    for (int d = d0; d < d0 + D_rep; ++d)
      st_cs_v4(vol + ((static_cast<size_t>(b) * Dvol + d) * plane + hw) * Cvol + ch_off + c8 * 8, o);
  }
}

}  // namespace

extern "C" int dpf_asm_sample_fwd(const void* x, void* out, int B, int H4, int W4, int C, int S, const int* ri,
                                  const float* rw, const int* ci, const float* cw, void* stream) {
  DPF_REQUIRE(x && out && ri && rw && ci && cw, "dpf_asm_sample_fwd: null pointer");
  DPF_REQUIRE(DPF_ALIGNED16(x) && DPF_ALIGNED16(out), "dpf_asm_sample_fwd: pointers must be 16-byte aligned");
  DPF_REQUIRE(C >= 8 && C % 8 == 0 && S >= 1 && S <= 8 && B > 0 && H4 > 0 && W4 > 0, "dpf_asm_sample_fwd: bad shape");
  const long long total = static_cast<long long>(B) * S * H4 * W4 * (C / 8);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(dpf::sm_count()) * 16));
  asm_sample_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), B, H4, W4, C, S, ri, rw, ci, cw);
  return dpf::after_launch("dpf_asm_sample_fwd");
}

extern "C" long long dpf_channel_stats_ws_floats(int B, long long P, int C) {
  if (B <= 0 || P <= 0 || C < 8 || C % 8) return 0;
  return static_cast<long long>(B) * channel_stats_chunks(B, P, C) * 2 * C;
}

extern "C" int dpf_channel_stats(const void* x, float* stats, float* ws, int B, long long P, int C, void* stream) {
  DPF_REQUIRE(x && stats && ws, "dpf_channel_stats: null pointer (ws = caller-owned workspace of dpf_channel_stats_ws_floats() floats)");
  DPF_REQUIRE(DPF_ALIGNED16(x), "dpf_channel_stats: x must be 16-byte aligned");
  DPF_REQUIRE(C >= 8 && C % 8 == 0 && C <= 256 && (256 % (C / 8)) == 0 && B > 0 && B <= 65535 && P > 0, "dpf_channel_stats: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int lanes = 256 / (C / 8);
  const int chunks = channel_stats_chunks(B, P, C);
  channel_stats_partial_kernel<<<dim3(chunks, B), 256, static_cast<size_t>(lanes) * 2 * C * sizeof(float), st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ws, P, C);
  if (int rc = dpf::after_launch("dpf_channel_stats")) return rc;
  channel_stats_final_kernel<<<B, 2 * C, 0, st>>>(ws, stats, chunks, C);
  return dpf::after_launch("dpf_channel_stats");
}

extern "C" int dpf_asm_blend_fwd(const void* samples, const void* logits, const float* in_a, const float* in_d, void* vol,
                                 int B, int H4, int W4, int C, int S, int D_vol, int d0, int D_rep, int ch_off, int Cvol,
                                 void* stream) {
  DPF_REQUIRE(samples && logits && in_a && in_d && vol, "dpf_asm_blend_fwd: null pointer");
  DPF_REQUIRE(DPF_ALIGNED16(samples) && DPF_ALIGNED16(logits) && DPF_ALIGNED16(vol), "dpf_asm_blend_fwd: pointers must be 16-byte aligned");
  DPF_REQUIRE(C >= 8 && C % 8 == 0 && ch_off % 8 == 0 && Cvol % 8 == 0 && ch_off + C <= Cvol, "dpf_asm_blend_fwd: bad channel layout");
  DPF_REQUIRE(S >= 1 && S <= 3, "dpf_asm_blend_fwd: S=%d must be 1..3", S);
  DPF_REQUIRE(D_rep >= 1 && d0 >= 0 && d0 + D_rep <= D_vol, "dpf_asm_blend_fwd: bad level range d0=%d D_rep=%d D_vol=%d", d0, D_rep, D_vol);
  const long long total = static_cast<long long>(B) * H4 * W4 * (C / 8);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(dpf::sm_count()) * 16));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto sp = reinterpret_cast<const __nv_bfloat16*>(samples);
  auto lg = reinterpret_cast<const __nv_bfloat16*>(logits);
  auto vo = reinterpret_cast<__nv_bfloat16*>(vol);
  if (S == 3) asm_blend_kernel<3><<<blocks, 256, 0, st>>>(sp, lg, in_a, in_d, vo, B, H4, W4, C, D_vol, d0, D_rep, ch_off, Cvol);
  else if (S == 2) asm_blend_kernel<2><<<blocks, 256, 0, st>>>(sp, lg, in_a, in_d, vo, B, H4, W4, C, D_vol, d0, D_rep, ch_off, Cvol);
  else asm_blend_kernel<1><<<blocks, 256, 0, st>>>(sp, lg, in_a, in_d, vo, B, H4, W4, C, D_vol, d0, D_rep, ch_off, Cvol);
  return dpf::after_launch("dpf_asm_blend_fwd");
}
