// Micro-benchmark: cost of a 128-byte-line gather as a function of how many distinct lines ONE load instruction touches.
// V = bytes per lane (4, 8, 16, 32); a line is covered by 128/V consecutive lanes, so a warp instruction touches 32*V/128 lines.
// Reports lines per clock per SM for global loads that hit L1 (64 KB footprint), L2 (32 MB footprint) and for shared memory.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int V> struct Vec;
template <> struct Vec<4> { using T = uint32_t; };
template <> struct Vec<8> { using T = uint2; };
template <> struct Vec<16> { using T = uint4; };
struct alignas(32) U8 { uint4 a, b; };
template <> struct Vec<32> { using T = U8; };

__device__ __forceinline__ uint32_t fold(uint32_t v) { return v; }
__device__ __forceinline__ uint32_t fold(uint2 v) { return v.x ^ v.y; }
__device__ __forceinline__ uint32_t fold(uint4 v) { return v.x ^ v.y ^ v.z ^ v.w; }
__device__ __forceinline__ uint32_t fold(U8 v) { return fold(v.a) ^ fold(v.b); }

template <int V> __device__ __forceinline__ typename Vec<V>::T ldv(const char* p) { return *reinterpret_cast<const typename Vec<V>::T*>(p); }
template <> __device__ __forceinline__ U8 ldv<32>(const char* p) {
  U8 r;
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w) : "l"(p));
  return r;
}

template <int V, bool SMEM>
__global__ void __launch_bounds__(512, 1) gather(const char* __restrict__ g, uint32_t nlines_mask, int iters, uint32_t* out, long long* cyc) {
  extern __shared__ __align__(128) char sm[];
  constexpr int L = 128 / V;                       // lanes per line
  const int lane = threadIdx.x & 31;
  const uint32_t grp = (blockIdx.x * blockDim.x + threadIdx.x) / L;      // one "voxel" per lane group
  const uint32_t sub = (lane % L) * V;
  if (SMEM) {
    for (int i = threadIdx.x * 16; i < 64 * 1024; i += blockDim.x * 16) *reinterpret_cast<uint4*>(sm + i) = *reinterpret_cast<const uint4*>(g + i);
    __syncthreads();
  }
  const char* base = SMEM ? sm : g;
  uint32_t s = grp * 2654435761u + 12345u, acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    typename Vec<V>::T r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {                  // 8 independent loads in flight (the 8 corners of a sample)
      s = s * 1664525u + 1013904223u;
      const uint32_t line = (s >> 8) & nlines_mask;
      if (SMEM) r[k] = *reinterpret_cast<const typename Vec<V>::T*>(base + line * 128u + sub);
      else r[k] = ldv<V>(base + static_cast<size_t>(line) * 128u + sub);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc ^= fold(r[k]);
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V, bool SMEM>
void run(const char* name, const char* g, uint32_t nlines, int iters, uint32_t* out, long long* cyc) {
  const int grid = 148;
  const int smem = SMEM ? 64 * 1024 : 0;
  cudaFuncSetAttribute(gather<V, SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  gather<V, SMEM><<<grid, 512, smem>>>(g, nlines - 1, 50, out, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  gather<V, SMEM><<<grid, 512, smem>>>(g, nlines - 1, iters, out, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  const double lines_per_cta = 512.0 / (128 / V) * 8.0 * iters;
  printf("%-22s V=%2d B/lane lines/instr=%2d : %.3f lines/clk/SM  (%.2f clk/line, %.2f clk/warp-instr)  %.3f ms  err=%s\n", name, V, 32 * V / 128,
         lines_per_cta / avg, avg / lines_per_cta, avg / (16.0 * 8.0 * iters), ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  char* g; uint32_t* out; long long* cyc;
  const size_t bytes = 64u << 20;
  cudaMalloc(&g, bytes); cudaMemset(g, 1, bytes); cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
  const int it = 2000;
  for (int pass = 0; pass < 2; ++pass) {
    const uint32_t nl = pass == 0 ? 512 : (32u << 20) / 128;            // 64 KB (L1-resident) / 32 MB (L2-resident)
    const char* nm = pass == 0 ? "LDG L1-hit (64 KB)" : "LDG L2-hit (32 MB)";
    run<4, false>(nm, g, nl, it, out, cyc); run<8, false>(nm, g, nl, it, out, cyc);
    run<16, false>(nm, g, nl, it, out, cyc); run<32, false>(nm, g, nl, it, out, cyc);
  }
  run<4, true>("LDS (64 KB)", g, 512, it, out, cyc); run<8, true>("LDS (64 KB)", g, 512, it, out, cyc); run<16, true>("LDS (64 KB)", g, 512, it, out, cyc);
  return 0;
}
